// stand-in that SHADOWS the reference header of the same name on the include path of the reference pin: the target-observation
// container pulls in OpenCV images, Boost smart pointers and serialisation; the camera headers only name a few of its accessors inside
// their initialisation members (parsed, never run by oracle/ref_pin.cpp).
#ifndef KB_SHIM_ASLAM_GRID_OBSERVATION
#define KB_SHIM_ASLAM_GRID_OBSERVATION
#include <Eigen/Core>
#include <memory>
#include <opencv2/core/core.hpp>
#include <vector>
namespace aslam { namespace cameras {
class GridCalibrationTargetBase {
 public:
  size_t rows() const { return 0; }
  size_t cols() const { return 0; }
  size_t size() const { return 0; }
  Eigen::Vector3d point(size_t) const { return Eigen::Vector3d(); }
};
class GridCalibrationTargetObservation {
 public:
  std::shared_ptr<const GridCalibrationTargetBase> target() const { return std::shared_ptr<const GridCalibrationTargetBase>(); }
  bool imagePoint(size_t, Eigen::Vector2d&) const { return false; }
  bool imageGridPoint(size_t, size_t, Eigen::Vector2d&) const { return false; }
  unsigned int getCornersImageFrame(std::vector<cv::Point2f>&) const { return 0; }
  unsigned int getCornersTargetFrame(std::vector<cv::Point3f>&) const { return 0; }
  unsigned int getCornersIdx(std::vector<unsigned int>&) const { return 0; }
  size_t imRows() const { return 0; }
  size_t imCols() const { return 0; }
  bool hasSuccessfulObservation() const { return false; }
};
} }
#endif
