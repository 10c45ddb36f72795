/* oracle/shim: stand-in for <boost/math/distributions/lognormal.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_MATH_DISTRIBUTIONS_LOGNORMAL_HPP
#define ORC_SHIM_BOOST_MATH_DISTRIBUTIONS_LOGNORMAL_HPP
#include <orc_boost_shim.hpp>
#endif
