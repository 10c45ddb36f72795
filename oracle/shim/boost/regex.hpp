/* oracle/shim: stand-in for <boost/regex.hpp> -- TEST INFRASTRUCTURE (see orc_graph_shim.hpp). */
#ifndef ORC_SHIM_BOOST_REGEX_HPP
#define ORC_SHIM_BOOST_REGEX_HPP
#include <orc_graph_shim.hpp>
#endif
