/* oracle/shim: stand-in for <boost/graph/adjacency_list.hpp> -- TEST INFRASTRUCTURE (see orc_graph_shim.hpp). */
#ifndef ORC_SHIM_BOOST_GRAPH_ADJACENCY_LIST_HPP
#define ORC_SHIM_BOOST_GRAPH_ADJACENCY_LIST_HPP
#include <orc_graph_shim.hpp>
#endif
