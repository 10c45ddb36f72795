/* oracle/shim: stand-in for <boost/algorithm/string.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_ALGORITHM_STRING_HPP
#define ORC_SHIM_BOOST_ALGORITHM_STRING_HPP
#include <orc_boost_shim.hpp>
#include <orc_graph_shim.hpp>
#endif
