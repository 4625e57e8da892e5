// stand-in: the OpenCV names the camera headers mention in their initialisation / PnP members (only parsed, never run here)
#ifndef KB_SHIM_OPENCV_CORE
#define KB_SHIM_OPENCV_CORE
#include <Eigen/Core>
#include <vector>
#define CV_64F 6
#define CV_32F 5
#define CV_64FC1 6
namespace cv {
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
  template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point_<double> Point2d;
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point3_<double> Point3d;
typedef Point3_<float> Point3f;
template <typename T> Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> double norm(const Point_<T>& a) { return std::sqrt((double)a.x * a.x + (double)a.y * a.y); }
class Mat {
 public:
  Mat() : rows(0), cols(0) {}
  Mat(int r, int c, int) : rows(r), cols(c), v((size_t)r * c, 0.0) {}
  static Mat eye(int r, int c, int t) { Mat m(r, c, t); for (int i = 0; i < r && i < c; ++i) m.v[(size_t)i * c + i] = 1.0; return m; }
  static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
  template <typename T> T& at(int r, int c = 0) { return *reinterpret_cast<T*>(&v[(size_t)r * cols + c]); }
  Mat t() const { return *this; }
  Mat inv() const { return *this; }
  Mat clone() const { return *this; }
  Mat rowRange(int, int) const { return *this; }
  Mat colRange(int, int) const { return *this; }
  int rows, cols;
  std::vector<double> v;
};
inline Mat operator*(const Mat& a, const Mat&) { return a; }
template <typename A, typename B, typename C, typename D, typename E, typename F>
bool solvePnP(const A&, const B&, const C&, const D&, E&, F&, bool = false, int = 0) { return false; }
template <typename A, typename B> void Rodrigues(const A&, B&) {}
template <typename A, typename B> void cv2eigen(const A&, B&) {}
template <typename A, typename B> void eigen2cv(const A&, B&) {}
struct SVD {
  enum { MODIFY_A = 1, FULL_UV = 4 };
  template <typename A, typename B> static void solveZ(const A&, B&) {}
};
}  // namespace cv
#endif
