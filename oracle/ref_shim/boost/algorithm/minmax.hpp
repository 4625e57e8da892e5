// stand-in: boost::minmax
#ifndef KB_SHIM_BOOST_MINMAX
#define KB_SHIM_BOOST_MINMAX
#include <utility>
namespace boost { template <typename T> std::pair<const T&, const T&> minmax(const T& a, const T& b) { return b < a ? std::pair<const T&, const T&>(b, a) : std::pair<const T&, const T&>(a, b); } }
#endif
