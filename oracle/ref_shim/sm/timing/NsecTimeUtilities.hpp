// stand-in: only included, not used, by the sources of the reference pin
#ifndef KB_SHIM_SM_NSEC
#define KB_SHIM_SM_NSEC
#include <cstdint>
namespace sm { namespace timing { typedef std::int64_t NsecTime; } }
#endif
