/* oracle/shim: stand-in for <boost/random/variate_generator.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_RANDOM_VARIATE_GENERATOR_HPP
#define ORC_SHIM_BOOST_RANDOM_VARIATE_GENERATOR_HPP
#include <orc_boost_shim.hpp>
#endif
