/* oracle/shim: stand-in for <boost/filesystem.hpp> -- TEST INFRASTRUCTURE (see orc_boost_shim.hpp).
 * fragmentLengthModel.cpp:156 only asks whether a file exists. */
#ifndef ORC_SHIM_BOOST_FILESYSTEM_HPP
#define ORC_SHIM_BOOST_FILESYSTEM_HPP
#include <fstream>
#include <string>
namespace boost { namespace filesystem {
inline bool exists(const std::string &p) { std::ifstream f(p.c_str()); return f.good(); }
} }
#endif
