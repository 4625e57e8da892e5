// stand-in that SHADOWS sm_eigen's matrix assertions (finite checks): no-ops in the reference pin
#ifndef KB_SHIM_SM_EIGEN_ASSERT_MACROS
#define KB_SHIM_SM_EIGEN_ASSERT_MACROS
#include <sm/assert_macros.hpp>
#define SM_ASSERT_MAT_IS_FINITE(exceptionType, matrix, message)
#define SM_ASSERT_MAT_IS_FINITE_DBG(exceptionType, matrix, message)
#endif
