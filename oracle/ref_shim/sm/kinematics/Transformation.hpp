// stand-in: just the members the camera headers name (only parsed, never run by oracle/ref_pin.cpp)
#ifndef KB_SHIM_SM_KINEMATICS_TRANSFORMATION
#define KB_SHIM_SM_KINEMATICS_TRANSFORMATION
#include <Eigen/Core>
namespace sm { namespace kinematics {
class Transformation {
 public:
  Transformation() : T_(Eigen::Matrix4d::Identity()) {}
  Transformation(const Eigen::Matrix4d& T) : T_(T) {}
  const Eigen::Matrix4d& T() const { return T_; }
  Eigen::Matrix3d C() const { return Eigen::Matrix3d(T_.topLeftCorner<3, 3>()); }
  Eigen::Vector3d t() const { return Eigen::Vector3d(T_.block<3, 1>(0, 3)); }
  Transformation inverse() const { return Transformation(Eigen::Matrix4d(T_.inverse())); }
  void set(const Eigen::Matrix4d& T) { T_ = T; }
  Transformation operator*(const Transformation& o) const { return Transformation(Eigen::Matrix4d(T_ * o.T_)); }
  Eigen::Vector3d operator*(const Eigen::Vector3d& p) const { return Eigen::Vector3d(C() * p + t()); }
 private:
  Eigen::Matrix4d T_;
};
} }
#endif
