#include <boost/thread.hpp>
