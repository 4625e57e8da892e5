// stand-in: CSparse's header is included by BE/include/aslam/backend/CompressedColumnMatrix.hpp:4, which uses nothing of it
#ifndef KB_SHIM_CS_H
#define KB_SHIM_CS_H
#endif
