// stand-in: class-version macros of sm_boost
#ifndef KB_SHIM_SM_BOOST_SERIALIZATION
#define KB_SHIM_SM_BOOST_SERIALIZATION
#include <boost/serialization/nvp.hpp>
#include <sm/assert_macros.hpp>
#define SM_BOOST_CLASS_VERSION(T)
#define SM_BOOST_CLASS_VERSION_T1(T)
#define SM_BOOST_CLASS_VERSION_T2(T)
#define SM_BOOST_CLASS_VERSION_T3(T)
#define SM_BOOST_CLASS_VERSION_I1(T)
#define SM_BOOST_CLASS_VERSION_T1I1(T)
#define SM_BOOST_CLASS_VERSION_T(T)
#endif
