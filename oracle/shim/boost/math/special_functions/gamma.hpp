/* oracle/shim: stand-in for <boost/math/special_functions/gamma.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_MATH_SPECIAL_FUNCTIONS_GAMMA_HPP
#define ORC_SHIM_BOOST_MATH_SPECIAL_FUNCTIONS_GAMMA_HPP
#include <orc_boost_shim.hpp>
#endif
