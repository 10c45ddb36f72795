/* oracle/shim: stand-in for <boost/graph/graphviz.hpp> -- TEST INFRASTRUCTURE (see orc_graph_shim.hpp). */
#ifndef ORC_SHIM_BOOST_GRAPH_GRAPHVIZ_HPP
#define ORC_SHIM_BOOST_GRAPH_GRAPHVIZ_HPP
#include <orc_graph_shim.hpp>
#endif
