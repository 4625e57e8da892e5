// stand-in: boost::math::round / isnan / isinf
#ifndef KB_SHIM_BOOST_MATH
#define KB_SHIM_BOOST_MATH
#include <cmath>
namespace boost { namespace math {
template <typename T> T round(T v) { return std::round(v); }
template <typename T> bool isnan(T v) { return std::isnan(v); }
template <typename T> bool isinf(T v) { return std::isinf(v); }
template <typename T> bool isfinite(T v) { return std::isfinite(v); }
} }
#endif
