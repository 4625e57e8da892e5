// stand-in: boost smart pointers -> the standard ones
#ifndef KB_SHIM_BOOST_SHARED_PTR
#define KB_SHIM_BOOST_SHARED_PTR
#include <memory>
#include <tuple>
namespace boost {
using std::shared_ptr;
using std::weak_ptr;
using std::make_shared;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
using std::const_pointer_cast;
using std::enable_shared_from_this;
using std::tie;
}  // namespace boost
#endif
