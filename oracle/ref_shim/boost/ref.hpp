#include <boost/function.hpp>
