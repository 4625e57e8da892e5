// stand-in that SHADOWS sm_eigen's Boost.Serialization adaptors for Eigen types: serialisation is not exercised by the reference pin
#ifndef KB_SHIM_SM_EIGEN_SERIALIZATION
#define KB_SHIM_SM_EIGEN_SERIALIZATION
#include <Eigen/Core>
#include <boost/serialization/nvp.hpp>
#endif
