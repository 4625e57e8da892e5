// stand-in: member-wise comparison helper of sm_common
#ifndef KB_SHIM_SM_SERIALIZATION_MACROS
#define KB_SHIM_SM_SERIALIZATION_MACROS
#define SM_CHECKMEMBERSSAME(other, member) ((other).member == this->member)
#define SM_CHECKSAME(a, b) ((a) == (b))
#endif
