// stand-in: boost::thread / thread_group / mutex -> the standard ones
#ifndef KB_SHIM_BOOST_THREAD
#define KB_SHIM_BOOST_THREAD
#include <boost/function.hpp>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
namespace boost {
class thread : public std::thread {
 public:
  using std::thread::thread;
  static unsigned hardware_concurrency() { return std::thread::hardware_concurrency(); }
};
class thread_group {
 public:
  ~thread_group() { join_all(); }
  template <typename F> thread* create_thread(F f) { t_.emplace_back(new thread(f)); return t_.back().get(); }
  void join_all() { for (auto& t : t_) if (t->joinable()) t->join(); }
 private:
  std::vector<std::unique_ptr<thread>> t_;
};
typedef std::mutex mutex;
template <typename M> using lock_guard = std::lock_guard<M>;
template <typename M> using unique_lock = std::unique_lock<M>;
}  // namespace boost
#endif
