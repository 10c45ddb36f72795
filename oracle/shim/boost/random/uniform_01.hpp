/* oracle/shim: stand-in for <boost/random/uniform_01.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_RANDOM_UNIFORM_01_HPP
#define ORC_SHIM_BOOST_RANDOM_UNIFORM_01_HPP
#include <orc_boost_shim.hpp>
#endif
