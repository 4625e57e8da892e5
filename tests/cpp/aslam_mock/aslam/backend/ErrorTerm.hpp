// forwards to the test stand-in of the reference header of the same path (tests/cpp/aslam_mock/aslam_backend_mock.hpp)
#include "../../aslam_backend_mock.hpp"
