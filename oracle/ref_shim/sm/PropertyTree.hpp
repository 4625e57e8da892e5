// stand-in: the configuration tree the reference's constructors can read from (never used by oracle/ref_pin.cpp)
#ifndef KB_SHIM_SM_PROPERTY_TREE
#define KB_SHIM_SM_PROPERTY_TREE
#include <string>
namespace sm {
class PropertyTree {
 public:
  double getDouble(const std::string&) const { return 0.0; }
  double getDouble(const std::string&, double d) const { return d; }
  int getInt(const std::string&) const { return 0; }
  int getInt(const std::string&, int d) const { return d; }
  bool getBool(const std::string&) const { return false; }
  bool getBool(const std::string&, bool d) const { return d; }
  std::string getString(const std::string&) const { return std::string(); }
  std::string getString(const std::string&, const std::string& d) const { return d; }
  bool doesKeyExist(const std::string&) const { return false; }
};
class ConstPropertyTree : public PropertyTree {
 public:
  ConstPropertyTree() {}
  ConstPropertyTree(const PropertyTree&, const std::string& = std::string()) {}
};
}  // namespace sm
#endif
