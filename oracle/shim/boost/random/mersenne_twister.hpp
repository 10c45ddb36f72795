/* oracle/shim: stand-in for <boost/random/mersenne_twister.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_RANDOM_MERSENNE_TWISTER_HPP
#define ORC_SHIM_BOOST_RANDOM_MERSENNE_TWISTER_HPP
#include <orc_boost_shim.hpp>
#endif
