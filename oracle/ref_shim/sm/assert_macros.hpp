// stand-in: the assertion macros of sm_common (throw the given exception type with the streamed message)
#ifndef KB_SHIM_SM_ASSERT_MACROS
#define KB_SHIM_SM_ASSERT_MACROS
#include <sstream>
#include <stdexcept>
#include <string>
#define SM_DEFINE_EXCEPTION(name, parent)                              \
  class name : public parent {                                          \
   public:                                                              \
    name(const char* m) : parent(m) {}                                  \
    name(const std::string& m) : parent(m) {}                           \
  };
#define SM_THROW(E, msg)                     \
  {                                           \
    std::stringstream sm_shim_ss;             \
    sm_shim_ss << msg;                        \
    throw E(sm_shim_ss.str());                \
  }
#define SM_ASSERT_TRUE(E, cond, msg) \
  if (!(cond)) SM_THROW(E, msg)
#define SM_ASSERT_FALSE(E, cond, msg) SM_ASSERT_TRUE(E, !(cond), msg)
#define SM_ASSERT_LE(E, a, b, msg) SM_ASSERT_TRUE(E, (a) <= (b), msg)
#define SM_ASSERT_LT(E, a, b, msg) SM_ASSERT_TRUE(E, (a) < (b), msg)
#define SM_ASSERT_GE(E, a, b, msg) SM_ASSERT_TRUE(E, (a) >= (b), msg)
#define SM_ASSERT_GT(E, a, b, msg) SM_ASSERT_TRUE(E, (a) > (b), msg)
#define SM_ASSERT_EQ(E, a, b, msg) SM_ASSERT_TRUE(E, (a) == (b), msg)
#define SM_ASSERT_NE(E, a, b, msg) SM_ASSERT_TRUE(E, (a) != (b), msg)
#define SM_ASSERT_NEAR(E, a, b, tol, msg) SM_ASSERT_TRUE(E, std::fabs((a) - (b)) <= (tol), msg)
#define SM_ASSERT_GE_LT(E, v, lo, hi, msg) SM_ASSERT_TRUE(E, (v) >= (lo) && (v) < (hi), msg)
#define SM_ASSERT_TRUE_DBG(E, cond, msg)
#define SM_ASSERT_FALSE_DBG(E, cond, msg)
#define SM_ASSERT_LE_DBG(E, a, b, msg)
#define SM_ASSERT_LT_DBG(E, a, b, msg)
#define SM_ASSERT_GE_DBG(E, a, b, msg)
#define SM_ASSERT_GT_DBG(E, a, b, msg)
#define SM_ASSERT_EQ_DBG(E, a, b, msg)
#define SM_ASSERT_NE_DBG(E, a, b, msg)
#define SM_ASSERT_NEAR_DBG(E, a, b, tol, msg)
#define SM_ASSERT_GE_LT_DBG(E, v, lo, hi, msg)
#endif
