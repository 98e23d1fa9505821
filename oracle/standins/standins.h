/*
 * standins.h -- TEST INFRASTRUCTURE ONLY (never compiled into the product).
 *
 * Stand-ins for the headers the reference's two ROS nodes include (ROS, OpenCV 2.4, cv_bridge,
 * image_transport, dynamic_reconfigure, popt, the generated jackal_nav messages), none of which exists in
 * this image.  With them src/obstacle_avoidance/point_cloud.cpp and navigate.cpp compile WHERE THEY LIE,
 * unmodified (oracle/Makefile, targets _ref/libpointcloud_ref.so and _ref/libnavigate_ref.so), and the
 * parity tests call the reference's own functions: cacheDisparityValues, publishObstacleScan (both),
 * publishPointCloud, generateDisparityMap, laserScanCallback, checkObstacle, chooseDirection.
 *
 * What this pins: the reference's STATEMENTS as its own compiler sees them -- loop order, constants, casts,
 * conditions, the order of min/max updates, message fields.  What it cannot pin: the arithmetic INSIDE
 * OpenCV, which is not under /root/reference.  The two OpenCV operations on the path are restated here and
 * pinned separately against OpenCV 4.13 (tests/golden/scan_cv2.npz):
 *   - Mat * Mat [+ Mat] on doubles: ((a0*b0 + a1*b1) + a2*b2) [+ a3*b3], then + C  (small-matrix gemm)
 *   - convertTo(CV_8U): saturate_cast<uchar>(cvRound(x)), round half to even
 * Everything else is a plain struct or a function that aborts if the tests ever reach it (FileStorage,
 * stereoRectify, initUndistortRectifyMap, imdecode, remap: the `main` and image callbacks of the nodes are
 * compiled but never run).  Messages handed to a ros::Publisher are kept in jn_standin::captured().
 * Nothing here is derived from ROS or OpenCV sources: only the names and signatures the two files use.
 */
#ifndef JN_ORACLE_STANDINS_H
#define JN_ORACLE_STANDINS_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <exception>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#define JN_STANDIN_UNREACHABLE(what)                                                     \
  do {                                                                                   \
    std::fprintf(stderr, "oracle stand-in: %s is not implemented (never on the tested path)\n", what); \
    std::abort();                                                                        \
  } while (0)

/* ------------------------------------------------------------------------------------------ OpenCV */
typedef unsigned char uchar;      /* OpenCV declares it in the global namespace too */

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC2 CV_MAKETYPE(CV_8U, 2)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_CALIB_ZERO_DISPARITY 1024

namespace cv {

template <class T, int N>
struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; i++) val[i] = T(); }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 2> Vec2b;
typedef Vec<uchar, 3> Vec3b;
typedef Vec<double, 3> Vec3d;

struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

template <class T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<double> Point2d;

template <class T>
struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<double> Point3d;

template <class T> struct DataType;
template <> struct DataType<double> { enum { type = CV_64FC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };

inline int jn_depth_bytes(int type) {
  switch (type & 7) {
    case CV_8U: return 1;
    case CV_32F: return 4;
    case CV_64F: return 8;
  }
  JN_STANDIN_UNREACHABLE("a cv::Mat depth other than 8U / 32F / 64F");
}

/* Dense row-major matrix with shared storage: enough of cv::Mat for the two files. */
class Mat {
 public:
  int rows, cols;
  int type_;
  size_t step;      /* bytes per row */
  uchar* data;
  std::shared_ptr<std::vector<uchar> > buf;

  Mat() : rows(0), cols(0), type_(0), step(0), data(0) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); fill(s); }
  Mat(Size sz, int type) { create(sz.height, sz.width, type); }
  Mat(Size sz, int type, const Scalar& s) { create(sz.height, sz.width, type); fill(s); }
  explicit Mat(const std::vector<uchar>& v) {            /* cv::Mat(msg->data): an n x 1 copy */
    create((int)v.size(), 1, CV_8UC1);
    if (!v.empty()) std::memcpy(data, &v[0], v.size());
  }
  explicit Mat(const Vec3d& v) {
    create(3, 1, CV_64FC1);
    for (int i = 0; i < 3; i++) at<double>(i, 0) = v[i];
  }

  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = (size_t)c * elemSize();
    buf.reset(new std::vector<uchar>(step * (size_t)r + 16, 0));
    data = r * c ? &(*buf)[0] : 0;
  }
  int channels() const { return (type_ >> 3) + 1; }
  int type() const { return type_; }
  size_t elemSize() const { return (size_t)jn_depth_bytes(type_) * channels(); }
  bool empty() const { return data == 0 || rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }

  void fill(const Scalar& s) {
    const int cn = channels();
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++)
        for (int k = 0; k < cn; k++) {
          uchar* p = data + r * step + ((size_t)c * cn + k) * jn_depth_bytes(type_);
          switch (type_ & 7) {
            case CV_8U: { double v = s.val[k]; *p = (uchar)(v < 0 ? 0 : v > 255 ? 255 : (int)lrint(v)); break; }
            case CV_32F: *(float*)p = (float)s.val[k]; break;
            case CV_64F: *(double*)p = s.val[k]; break;
          }
        }
  }
  static Mat eye(int r, int c, int type) {
    Mat m(r, c, type, Scalar(0));
    if (type != CV_64FC1) JN_STANDIN_UNREACHABLE("Mat::eye of a type other than CV_64FC1");
    for (int i = 0; i < std::min(r, c); i++) m.at<double>(i, i) = 1.0;
    return m;
  }
  static Mat zeros(Size sz, int type) { return Mat(sz, type, Scalar(0)); }

  template <class T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> T* ptr(int r) { return (T*)(data + (size_t)r * step); }

  Mat operator()(const Rect& roi) const {                /* a view: same storage, same step */
    Mat m(*this);
    m.rows = roi.height; m.cols = roi.width;
    m.data = data + (size_t)roi.y * step + (size_t)roi.x * elemSize();
    return m;
  }

  /* convertTo(dst, CV_8U, 1.) of a float map: saturate_cast<uchar>(cvRound(x)) -- round half to even, then clamp.
   * Pinned against cv2 4.13 in tests/golden/scan_cv2.npz. */
  void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
    if ((type_ & 7) != CV_32F || channels() != 1 || (rtype & 7) != CV_8U || alpha != 1.0 || beta != 0.0)
      JN_STANDIN_UNREACHABLE("convertTo other than CV_32F -> CV_8U with alpha 1, beta 0");
    if (dst.rows != rows || dst.cols != cols || dst.type_ != CV_8UC1) dst.create(rows, cols, CV_8UC1);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        const int v = (int)lrintf(at<float>(r, c));
        dst.at<uchar>(r, c) = (uchar)(v < 0 ? 0 : v > 255 ? 255 : v);
      }
  }

  template <class T> operator Point3_<T>() const {       /* Point3d(point3d_robot) */
    if (type_ != CV_64FC1 || rows * cols != 3) JN_STANDIN_UNREACHABLE("Mat -> Point3 of a shape other than 3 doubles");
    return Point3_<T>((T)at<double>(0, 0), (T)((const double*)data)[1], (T)((const double*)data)[2]);
  }
};

/* d = a * b [+ c] on CV_64FC1, the summation order of OpenCV's small-matrix gemm. */
inline Mat jn_gemm(const Mat& a, const Mat& b, const Mat* c) {
  if (a.type_ != CV_64FC1 || b.type_ != CV_64FC1 || a.cols != b.rows || (c && (c->rows != a.rows || c->cols != b.cols)))
    JN_STANDIN_UNREACHABLE("a matrix product other than CV_64FC1 with matching shapes");
  Mat d(a.rows, b.cols, CV_64FC1);
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < b.cols; j++) {
      double s = a.at<double>(i, 0) * b.at<double>(0, j);
      for (int k = 1; k < a.cols; k++) s = s + a.at<double>(i, k) * b.at<double>(k, j);
      d.at<double>(i, j) = c ? s + c->at<double>(i, j) : s;
    }
  return d;
}

struct MatMulExpr {            /* what `A * B` is before it is assigned or added to */
  Mat a, b;
  operator Mat() const { return jn_gemm(a, b, 0); }
};
inline MatMulExpr operator*(const Mat& a, const Mat& b) { MatMulExpr e; e.a = a; e.b = b; return e; }
inline MatMulExpr operator*(const MatMulExpr& l, const Mat& b) { MatMulExpr e; e.a = Mat(l); e.b = b; return e; }
inline Mat operator+(const MatMulExpr& e, const Mat& c) { return jn_gemm(e.a, e.b, &c); }

template <class T>
struct Mat_ : public Mat {
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
};
template <class T>
struct MatCommaInitializer_ {  /* (Mat_<double>(3,1) << x, y, z) */
  Mat m;
  int n;
  MatCommaInitializer_& operator,(T v) { ((T*)m.data)[n++] = v; return *this; }
  operator Mat() const { return m; }
};
template <class T, class V>
inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v) {
  MatCommaInitializer_<T> ci;
  ci.m = m; ci.n = 0;
  return (ci, (T)v);
}
inline std::ostream& operator<<(std::ostream& os, const Mat& m) {
  os << "[";
  for (int r = 0; r < m.rows; r++)
    for (int c = 0; c < m.cols; c++)
      if (m.type_ == CV_64FC1) os << m.at<double>(r, c) << (c + 1 < m.cols ? ", " : r + 1 < m.rows ? ";\n " : "");
  return os << "]";
}

/* ---- compiled, never run by the tests */
enum { INTER_LINEAR = 1 };
struct FileNode {
  template <class T> void operator>>(T&) const { JN_STANDIN_UNREACHABLE("cv::FileStorage"); }
};
struct FileStorage {
  enum { READ = 0 };
  FileStorage(const char*, int) {}
  FileNode operator[](const char*) const { return FileNode(); }
};
inline Mat imdecode(const Mat&, int) { JN_STANDIN_UNREACHABLE("cv::imdecode"); }
inline void remap(const Mat&, Mat&, const Mat&, const Mat&, int) { JN_STANDIN_UNREACHABLE("cv::remap"); }
inline void stereoRectify(const Mat&, const Mat&, const Mat&, const Mat&, Size, const Mat&, const Mat&, Mat&, Mat&, Mat&,
                          Mat&, Mat&, int, double, Size, Rect*, Rect*) { JN_STANDIN_UNREACHABLE("cv::stereoRectify"); }
inline void initUndistortRectifyMap(const Mat&, const Mat&, const Mat&, const Mat&, Size, int, Mat&, Mat&) {
  JN_STANDIN_UNREACHABLE("cv::initUndistortRectifyMap");
}

}  // namespace cv

/* --------------------------------------------------------------------------------------------- ROS */
namespace ros {
struct Time {
  static Time now() { return Time(); }
};
}  // namespace ros

namespace std_msgs {
struct Header {
  uint32_t seq;
  ros::Time stamp;
  std::string frame_id;
  Header() : seq(0) {}
};
}  // namespace std_msgs

namespace geometry_msgs {
struct Point { double x, y, z; Point() : x(0), y(0), z(0) {} };
struct Point32 { float x, y, z; Point32() : x(0), y(0), z(0) {} };
struct Quaternion { double x, y, z, w; Quaternion() : x(0), y(0), z(0), w(0) {} };
struct Vector3 { double x, y, z; Vector3() : x(0), y(0), z(0) {} };
struct Pose { Point position; Quaternion orientation; };
struct Twist { Vector3 linear, angular; };
}  // namespace geometry_msgs

namespace sensor_msgs {
struct LaserScan {
  std_msgs::Header header;
  float angle_min, angle_max, angle_increment, time_increment, scan_time, range_min, range_max;
  std::vector<float> ranges, intensities;
  LaserScan() : angle_min(0), angle_max(0), angle_increment(0), time_increment(0), scan_time(0), range_min(0), range_max(0) {}
};
typedef std::shared_ptr<const LaserScan> LaserScanConstPtr;
struct ChannelFloat32 {
  std::string name;
  std::vector<float> values;
};
struct PointCloud {
  std_msgs::Header header;
  std::vector<geometry_msgs::Point32> points;
  std::vector<ChannelFloat32> channels;
};
struct CompressedImage {
  std_msgs::Header header;
  std::string format;
  std::vector<uint8_t> data;
};
typedef std::shared_ptr<const CompressedImage> CompressedImageConstPtr;
struct Image { std_msgs::Header header; };
typedef std::shared_ptr<Image> ImagePtr;
struct Joy {
  std_msgs::Header header;
  std::vector<float> axes;
  std::vector<int32_t> buttons;
};
typedef std::shared_ptr<const Joy> JoyConstPtr;
}  // namespace sensor_msgs

namespace visualization_msgs {
struct ColorRGBA { float r, g, b, a; ColorRGBA() : r(0), g(0), b(0), a(0) {} };
struct Marker {
  enum { ADD = 0, POINTS = 8 };
  std_msgs::Header header;
  std::string ns;
  int32_t id, type, action;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  ColorRGBA color;
  std::vector<geometry_msgs::Point> points;
  Marker() : id(0), type(0), action(0) {}
};
}  // namespace visualization_msgs

namespace jackal_nav {
struct JackalPose { double x, y, theta; JackalPose() : x(0), y(0), theta(0) {} };
typedef std::shared_ptr<const JackalPose> JackalPoseConstPtr;
struct JackalTimeLog {
  std_msgs::Header header;
  float pcl_time, obstacle_scan_time, dmap_time;
  JackalTimeLog() : pcl_time(0), obstacle_scan_time(0), dmap_time(0) {}
};
struct CamToRobotCalibParamsConfig {       /* cfg/CamToRobotCalibParams.cfg: six doubles with these defaults */
  double PHI_X, PHI_Y, PHI_Z, TRANS_X, TRANS_Y, TRANS_Z;
  CamToRobotCalibParamsConfig() : PHI_X(1.3), PHI_Y(-3.14), PHI_Z(1.57), TRANS_X(0), TRANS_Y(0), TRANS_Z(0.28) {}
};
}  // namespace jackal_nav

/* What the nodes publish is kept for the tests. */
namespace jn_standin {
struct Captured {
  std::vector<sensor_msgs::LaserScan> scans;
  std::vector<sensor_msgs::PointCloud> clouds;
  std::vector<geometry_msgs::Twist> twists;
  long others;
  Captured() : others(0) {}
};
inline Captured& captured() { static Captured c; return c; }
template <class M> inline void capture(const M&) { captured().others++; }
inline void capture(const sensor_msgs::LaserScan& m) { captured().scans.push_back(m); }
inline void capture(const sensor_msgs::PointCloud& m) { captured().clouds.push_back(m); }
inline void capture(const geometry_msgs::Twist& m) { captured().twists.push_back(m); }
}  // namespace jn_standin

namespace ros {
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
struct Publisher {
  template <class M> void publish(const M& m) const { jn_standin::capture(m); }
};
struct Subscriber {};
struct NodeHandle {
  template <class M> Publisher advertise(const std::string&, int) { return Publisher(); }
  template <class F> Subscriber subscribe(const std::string&, int, F) { return Subscriber(); }
};
}  // namespace ros

namespace image_transport {
struct Publisher {
  void publish(const sensor_msgs::ImagePtr&) const { jn_standin::captured().others++; }
};
struct ImageTransport {
  explicit ImageTransport(const ros::NodeHandle&) {}
  Publisher advertise(const std::string&, int) { return Publisher(); }
};
}  // namespace image_transport

namespace cv_bridge {
struct Exception : public std::exception {};
struct CvImage {
  CvImage(const std_msgs::Header&, const std::string&, const cv::Mat&) {}
  sensor_msgs::ImagePtr toImageMsg() const { return sensor_msgs::ImagePtr(new sensor_msgs::Image()); }
};
}  // namespace cv_bridge

namespace dynamic_reconfigure {
template <class C>
struct Server {
  typedef std::function<void(C&, uint32_t)> CallbackType;
  void setCallback(const CallbackType&) {}
};
}  // namespace dynamic_reconfigure

namespace boost {
template <class F, class A, class B>
inline auto bind(F f, A a, B b) -> decltype(std::bind(f, a, b)) { return std::bind(f, a, b); }
}  // namespace boost
using std::placeholders::_1;
using std::placeholders::_2;

/* -------------------------------------------------------------------------------------------- popt */
struct poptOption {
  const char* longName;
  char shortName;
  unsigned int argInfo;
  void* arg;
  int val;
  const char* descrip;
  const char* argDescrip;
};
typedef struct jn_popt_context* poptContext;
#define POPT_ARG_NONE 0U
#define POPT_ARG_STRING 1U
#define POPT_ARG_INT 2U
#define POPT_ARG_FLOAT 8U
#define POPT_BADOPTION_NOALIAS 1
#define POPT_AUTOHELP { NULL, '\0', 0U, NULL, 0, "Help options:", NULL },
inline poptContext poptGetContext(const char*, int, const char**, const poptOption*, unsigned int) { return 0; }
inline poptContext poptFreeContext(poptContext) { return 0; }
inline int poptGetNextOpt(poptContext) { return -1; }
inline const char* poptGetOptArg(poptContext) { return 0; }
inline const char* poptStrerror(int) { return ""; }
inline const char* poptBadOption(poptContext, unsigned int) { return ""; }
inline const char* poptGetArg(poptContext) { return 0; }

#endif /* JN_ORACLE_STANDINS_H */
