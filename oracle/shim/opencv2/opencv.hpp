// ORACLE — TEST INFRASTRUCTURE ONLY. Never included, linked or executed by the product path (hso_b200/).
//
// Stand-in for the OpenCV C++ headers the reference is written against (system dependency, version unpinned, CMakeLists.txt:44-50; absent from
// this image — only the Python cv2 4.13 wheel exists). It provides cv::Mat (a reference-counted byte buffer with OpenCV's public fields) and
// the handful of cv:: functions the hot-path translation units call. Their arithmetic is NOT invented here: each one forwards to the oracle's
// restatement of the published OpenCV algorithm, which tests/test_oracle_pins.py / test_oracle_undistort.py / test_oracle_reproject.py pin
// bit-for-bit to golden vectors generated with cv2 4.13 (tests/golden/make_cv_golden*.py):
//   cv::resize(INTER_LINEAR, 8UC1)  -> orc_resize_linear_u8        cv::Sobel(CV_16S, k = 5, BORDER_REPLICATE) -> orc_sobel5
//   cv::remap(INTER_LINEAR, CV_16SC2 maps) -> orc_remap_linear_u8  cv::convertMaps -> orc_convert_maps
//   cv::initUndistortRectifyMap(CV_16SC2)  -> orc_cv_init_undistort_rectify_map   cv::undistortPoints (1 point, 5 iterations) -> orc_cv_undistort_point
// GUI / debugging calls (imshow, namedWindow, waitKey) are no-ops.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)

#define CV_Assert(expr) assert(expr)
#define CV_DbgAssert(expr) assert(expr)
inline int cvRound(double v) { return (int)std::nearbyint(v); }  // round-half-to-even under the default rounding mode, like lrint / cvtsd2si
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { WINDOW_NORMAL = 0, WINDOW_AUTOSIZE = 1 };

template <class T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  T area() const { return width * height; }
  bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size;
template <class T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T xx, T yy) : x(xx), y(yy) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <class T, int N> struct Vec {
  T val[N];
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<short, 2> Vec2s;
typedef Vec<float, 2> Vec2f;
typedef Vec<uchar, 3> Vec3b;
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  double operator[](int i) const { return val[i]; }
};
struct Range { int start, end; Range(int s = 0, int e = 0) : start(s), end(e) {} };
struct Rect { int x, y, width, height; Rect(int a = 0, int b = 0, int c = 0, int d = 0) : x(a), y(b), width(c), height(d) {} };
struct KeyPoint {
  Point2f pt; float size, angle, response; int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
struct TermCriteria { int type, maxCount; double epsilon; TermCriteria(int t = 0, int m = 0, double e = 0) : type(t), maxCount(m), epsilon(e) {} enum { COUNT = 1, MAX_ITER = 1, EPS = 2 }; };

class Mat {
 public:
  struct MStep {
    size_t p[2];
    MStep() { p[0] = p[1] = 0; }
    operator size_t() const { return p[0]; }
    size_t operator[](int i) const { return p[i]; }
  };
  int flags, dims, rows, cols;
  uchar* data;
  MStep step;

  Mat() : flags(0), dims(2), rows(0), cols(0), data(nullptr) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar& v) : Mat() { create(r, c, type); setTo(v); }
  Mat(int r, int c, int type, void* ext, size_t stp = 0) : flags(type), dims(2), rows(r), cols(c), data((uchar*)ext) {  // user data: not owned, not copied
    step.p[1] = elemSize();
    step.p[0] = stp ? stp : (size_t)c * elemSize();
  }
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == this->type() && buf_) return;
    flags = type; rows = r; cols = c;
    step.p[1] = elemSize();
    step.p[0] = (size_t)c * elemSize();
    // 64-byte aligned like OpenCV's fastMalloc (the reference's halfSample takes its SSE2 path only for 16-byte aligned rows). Four zeroed rows
    // follow the image: CoarseTracker's forward-mode gradient taps read image row == rows for patches at the lower border
    // (src/CoarseTracker.cpp:370 with the :310 bounds test) — undefined in the reference (heap bytes), defined as 0 here, in the oracle
    // restatement and on the device alike.
    const size_t bytes = step.p[0] * ((size_t)r + 4) + 64;
    buf_ = std::shared_ptr<uchar>(new uchar[bytes + 64], std::default_delete<uchar[]>());
    data = (uchar*)(((uintptr_t)buf_.get() + 63) & ~(uintptr_t)63);
    std::memset(data, 0, bytes);
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  void release() { buf_.reset(); data = nullptr; rows = cols = 0; }
  int type() const { return flags & 0xFFF; }
  int depth() const { return CV_MAT_DEPTH(flags); }
  int channels() const { return CV_MAT_CN(flags); }
  size_t elemSize1() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)sz[depth()]; }
  size_t elemSize() const { return elemSize1() * (size_t)channels(); }
  size_t total() const { return (size_t)rows * cols; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step.p[0] == (size_t)cols * elemSize(); }
  Size size() const { return Size(cols, rows); }
  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type());
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step.p[0], data + (size_t)y * step.p[0], (size_t)cols * elemSize());
    return m;
  }
  void copyTo(Mat& dst) const { dst = clone(); }
  Mat& setTo(const Scalar& v) {
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols * channels(); ++x) {
        uchar* p = data + (size_t)y * step.p[0] + (size_t)x * elemSize1();
        const double d = v.val[x % channels()];
        switch (depth()) {
          case CV_8U: *p = (uchar)d; break;
          case CV_16S: *(short*)p = (short)d; break;
          case CV_16U: *(ushort*)p = (ushort)d; break;
          case CV_32S: *(int*)p = (int)d; break;
          case CV_32F: *(float*)p = (float)d; break;
          case CV_64F: *(double*)p = d; break;
          default: break;
        }
      }
    return *this;
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  static Mat zeros(Size s, int type) { return Mat(s, type); }
  template <class T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step.p[0] + (size_t)x * sizeof(T)); }
  template <class T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step.p[0] + (size_t)x * sizeof(T)); }
  template <class T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  template <class T> T& at(Point p) { return at<T>(p.y, p.x); }
  template <class T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.p[0]); }
  template <class T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.p[0]); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step.p[0]; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step.p[0]; }

 protected:
  std::shared_ptr<uchar> buf_;
};

template <class T> struct DataType;
template <> struct DataType<uchar> { enum { depth = CV_8U, type = CV_8UC1 }; };
template <> struct DataType<short> { enum { depth = CV_16S, type = CV_16SC1 }; };
template <> struct DataType<int> { enum { depth = CV_32S, type = CV_MAKETYPE(CV_32S, 1) }; };
template <> struct DataType<float> { enum { depth = CV_32F, type = CV_32FC1 }; };
template <> struct DataType<double> { enum { depth = CV_64F, type = CV_64FC1 }; };

inline size_t alignSize(size_t sz, int n) { return (sz + n - 1) & -n; }
template <class T> inline T* alignPtr(T* ptr, int n = (int)sizeof(T)) { return (T*)(((size_t)ptr + n - 1) & -n); }
template <class T> class AutoBuffer {
  std::vector<T> v_;
 public:
  explicit AutoBuffer(size_t n = 0) : v_(n + 8) {}
  operator T*() { return v_.data(); }
  operator const T*() const { return v_.data(); }
  size_t size() const { return v_.size(); }
};

template <class T> class Mat_;
template <class T> struct MatCommaInit_ {
  Mat_<T>* m; int k;
  template <class S> MatCommaInit_& operator,(S v);
  operator Mat_<T>() const;
};
template <class T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  T& operator()(int y, int x) { return this->template at<T>(y, x); }
  const T& operator()(int y, int x) const { return this->template at<T>(y, x); }
  static Mat_ eye(int r, int c) { Mat_ m(r, c); for (int i = 0; i < (r < c ? r : c); ++i) m(i, i) = T(1); return m; }
  template <class S> MatCommaInit_<T> operator<<(S v) { MatCommaInit_<T> ci{this, 0}; ci, v; return ci; }
};
template <class T> template <class S> MatCommaInit_<T>& MatCommaInit_<T>::operator,(S v) { (*m)(k / m->cols, k % m->cols) = (T)v; ++k; return *this; }
template <class T> MatCommaInit_<T>::operator Mat_<T>() const { return *m; }
// (cv::Mat_<float>(3,3) << a, b, ...) is assigned to a cv::Mat in the reference: the temporary's buffer is shared, not dangling
template <class T> inline Mat mat_from_init(const MatCommaInit_<T>& ci) { return (Mat)(*ci.m); }

typedef const Mat& InputArray;
typedef Mat& OutputArray;

// ---- the functions the hot-path TUs call; defined in oracle/shim/cv_shim.cpp on top of the oracle's cv2-pinned restatements -------------
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT);
void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interpolation, int borderMode = BORDER_CONSTANT, const Scalar& v = Scalar());
void convertMaps(const Mat& map1, const Mat& map2, Mat& dstmap1, Mat& dstmap2, int dstmap1type, bool nninterpolation = false);
void initUndistortRectifyMap(const Mat& K, const Mat& D, const Mat& R, const Mat& newK, Size size, int m1type, Mat& map1, Mat& map2);
void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D);
inline void split(const Mat&, std::vector<Mat>&) {}
inline void imshow(const std::string&, const Mat&) {}
inline void namedWindow(const std::string&, int = 0) {}
inline int waitKey(int = 0) { return -1; }
inline void destroyAllWindows() {}

}  // namespace cv
