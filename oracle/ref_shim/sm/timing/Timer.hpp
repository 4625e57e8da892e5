// stand-in that SHADOWS sm_timing's Timer (Boost.DateTime based): timers of the optimiser are no-ops in the reference pin
#ifndef KB_SHIM_SM_TIMING
#define KB_SHIM_SM_TIMING
#include <cstddef>
#include <string>
namespace sm { namespace timing {
class DummyTimer {
 public:
  DummyTimer(size_t = 0, bool = false) {}
  DummyTimer(std::string const&, bool = false) {}
  void start() {}
  void stop() {}
  void discardTiming() {}
  bool isTiming() const { return false; }
};
typedef DummyTimer Timer;
struct Timing {
  static size_t getHandle(std::string const&) { return 0; }
  static std::string print() { return std::string(); }
  template <typename OS> static void print(OS&) {}
  static void reset(std::string const&) {}
};
} }
#endif
