/* oracle/shim: stand-in for <boost/random/gamma_distribution.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_RANDOM_GAMMA_DISTRIBUTION_HPP
#define ORC_SHIM_BOOST_RANDOM_GAMMA_DISTRIBUTION_HPP
#include <orc_boost_shim.hpp>
#endif
