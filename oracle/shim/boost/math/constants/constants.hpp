/* oracle/shim: stand-in for <boost/math/constants/constants.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_MATH_CONSTANTS_CONSTANTS_HPP
#define ORC_SHIM_BOOST_MATH_CONSTANTS_CONSTANTS_HPP
#include <orc_boost_shim.hpp>
#endif
