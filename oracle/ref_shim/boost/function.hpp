// stand-in: boost::function / boost::bind / boost::ref -> the standard ones
#ifndef KB_SHIM_BOOST_FUNCTION
#define KB_SHIM_BOOST_FUNCTION
#include <functional>
namespace boost {
template <typename F> using function = std::function<F>;
using std::bind;
using std::ref;
using std::cref;
namespace placeholders { using namespace std::placeholders; }
}  // namespace boost
using namespace std::placeholders;
#endif
