// stand-in: logging macros of sm_logging (silent)
#ifndef KB_SHIM_SM_LOGGING
#define KB_SHIM_SM_LOGGING
#include <sstream>
#define SM_SHIM_LOG(x) do { std::stringstream sm_shim_log; sm_shim_log << x; } while (0)
#define SM_DEBUG_STREAM(x) SM_SHIM_LOG(x)
#define SM_INFO_STREAM(x) SM_SHIM_LOG(x)
#define SM_WARN_STREAM(x) SM_SHIM_LOG(x)
#define SM_ERROR_STREAM(x) SM_SHIM_LOG(x)
#define SM_FATAL_STREAM(x) SM_SHIM_LOG(x)
#define SM_DEBUG(...)
#define SM_INFO(...)
#define SM_WARN(...)
#define SM_ERROR(...)
#define SM_FATAL(...)
#endif
