/* oracle/shim: stand-in for <boost/property_map/property_map.hpp> -- TEST INFRASTRUCTURE (see orc_graph_shim.hpp). */
#ifndef ORC_SHIM_BOOST_PROPERTY_MAP_PROPERTY_MAP_HPP
#define ORC_SHIM_BOOST_PROPERTY_MAP_PROPERTY_MAP_HPP
#include <orc_graph_shim.hpp>
#endif
