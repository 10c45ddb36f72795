/* oracle/shim: stand-in for <boost/random/uniform_int_distribution.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_RANDOM_UNIFORM_INT_DISTRIBUTION_HPP
#define ORC_SHIM_BOOST_RANDOM_UNIFORM_INT_DISTRIBUTION_HPP
#include <orc_boost_shim.hpp>
#endif
