// stand-in: boost fixed-width integers
#ifndef KB_SHIM_BOOST_CSTDINT
#define KB_SHIM_BOOST_CSTDINT
#include <cstdint>
namespace boost { using std::uint64_t; using std::int64_t; using std::uint32_t; using std::int32_t; using std::uint8_t; using std::uint16_t; }
#endif
