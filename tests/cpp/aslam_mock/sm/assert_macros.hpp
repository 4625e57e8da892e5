// forwards to the test stand-in (tests/cpp/aslam_mock/aslam_backend_mock.hpp)
#include "../aslam_backend_mock.hpp"
