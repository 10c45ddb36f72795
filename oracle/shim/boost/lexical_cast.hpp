/* oracle/shim: stand-in for <boost/lexical_cast.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp). */
#ifndef ORC_SHIM_BOOST_LEXICAL_CAST_HPP
#define ORC_SHIM_BOOST_LEXICAL_CAST_HPP
#include <orc_boost_shim.hpp>
#endif
