/* oracle/shim: stand-in for <boost/graph/transitive_closure.hpp> -- TEST INFRASTRUCTURE (see orc_graph_shim.hpp). */
#ifndef ORC_SHIM_BOOST_GRAPH_TRANSITIVE_CLOSURE_HPP
#define ORC_SHIM_BOOST_GRAPH_TRANSITIVE_CLOSURE_HPP
#include <orc_graph_shim.hpp>
#endif
