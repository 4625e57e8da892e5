// stand-in: boost::hash -> std::hash
#ifndef KB_SHIM_BOOST_HASH
#define KB_SHIM_BOOST_HASH
#include <functional>
namespace boost { template <typename T> struct hash : std::hash<T> {}; }
#endif
