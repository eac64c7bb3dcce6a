// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or executed by the product path (hso_b200/).
//
// The cv:: functions declared in oracle/shim/opencv2/opencv.hpp, for the reference translation units compiled into oracle/_ref/libhso_ref.so.
// No OpenCV arithmetic is invented here: every function forwards to the oracle's restatement of the published OpenCV algorithm
// (oracle/oracle_frame.cpp, oracle_undistort.cpp, oracle_reproject.cpp), each of which is pinned bit-for-bit to golden vectors generated with
// cv2 4.13 (tests/golden/cv_golden.npz, cv_golden2.npz; tests/test_oracle_pins.py, test_oracle_undistort.py, test_oracle_reproject.py).
// Only the argument combinations the reference uses are supported; anything else aborts loudly.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <opencv2/opencv.hpp>

#include "../hso_oracle.h"

namespace {
[[noreturn]] void unsupported(const char* what) {
  std::fprintf(stderr, "oracle/shim/cv_shim.cpp: unsupported use of %s (only the reference's own call patterns are implemented)\n", what);
  std::abort();
}
// tightly packed copy of a CV_8UC1 matrix (the oracle functions take stride == cols)
std::vector<uint8_t> packed_u8(const cv::Mat& m) {
  std::vector<uint8_t> v((size_t)m.rows * m.cols);
  for (int y = 0; y < m.rows; ++y) std::memcpy(v.data() + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
  return v;
}
}  // namespace

namespace cv {

// frame.cpp:311, ImageReader.cpp:80: cv::resize(src, dst, Size, 0, 0, INTER_LINEAR) on CV_8UC1
void resize(const Mat& src, Mat& dst, Size dsize, double fx, double fy, int interpolation) {
  if (src.type() != CV_8UC1 || interpolation != INTER_LINEAR || fx != 0 || fy != 0 || dsize.width <= 0 || dsize.height <= 0) unsupported("cv::resize");
  const std::vector<uint8_t> s = packed_u8(src);  // also makes src == dst safe
  Mat out(dsize.height, dsize.width, CV_8UC1);
  orc_resize_linear_u8(s.data(), src.cols, src.rows, out.data, dsize.width, dsize.height);
  dst = out;
}

// frame.cpp:218-219: cv::Sobel(img, dst, CV_16S, 1, 0, 5, 1, 0, BORDER_REPLICATE) and (0, 1)
void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize, double scale, double delta, int borderType) {
  if (src.type() != CV_8UC1 || ddepth != CV_16S || ksize != 5 || scale != 1 || delta != 0 || borderType != BORDER_REPLICATE || dx + dy != 1)
    unsupported("cv::Sobel");
  const std::vector<uint8_t> s = packed_u8(src);
  std::vector<int16_t> gx((size_t)src.rows * src.cols), gy((size_t)src.rows * src.cols);
  orc_sobel5(s.data(), src.cols, src.rows, gx.data(), gy.data());
  Mat out(src.rows, src.cols, CV_16SC1);
  std::memcpy(out.data, dx == 1 ? gx.data() : gy.data(), gx.size() * sizeof(int16_t));
  dst = out;
}

// camera.cpp:130,270,369: cv::remap(raw, rectified, map1 (CV_16SC2), map2 (CV_16UC1), INTER_LINEAR)
void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interpolation, int borderMode, const Scalar&) {
  if (src.type() != CV_8UC1 || map1.type() != CV_16SC2 || map2.type() != CV_16UC1 || interpolation != INTER_LINEAR || borderMode != BORDER_CONSTANT)
    unsupported("cv::remap");
  const std::vector<uint8_t> s = packed_u8(src);
  Mat out(map1.rows, map1.cols, CV_8UC1);
  orc_remap_linear_u8(s.data(), src.cols, src.rows, src.cols, (const int16_t*)map1.data, (const uint16_t*)map2.data, map1.cols, map1.rows, out.data);
  dst = out;
}

// camera.cpp:244,339: cv::convertMaps(map_x_float, map_y_float, map1, map2, CV_16SC2)
void convertMaps(const Mat& map1, const Mat& map2, Mat& dstmap1, Mat& dstmap2, int dstmap1type, bool nninterpolation) {
  if (map1.type() != CV_32FC1 || map2.type() != CV_32FC1 || dstmap1type != CV_16SC2 || nninterpolation) unsupported("cv::convertMaps");
  Mat o1(map1.rows, map1.cols, CV_16SC2), o2(map1.rows, map1.cols, CV_16UC1);
  orc_convert_maps((const float*)map1.data, (const float*)map2.data, map1.rows * map1.cols, (int16_t*)o1.data, (uint16_t*)o2.data);
  dstmap1 = o1;
  dstmap2 = o2;
}

// camera.cpp:47-54: cv::initUndistortRectifyMap(cvK_ (float 3x3), cvD_ (float 1x5), eye(3,3), cvK_, size, CV_16SC2, map1, map2)
void initUndistortRectifyMap(const Mat& K, const Mat& D, const Mat& R, const Mat& newK, Size size, int m1type, Mat& map1, Mat& map2) {
  if (K.type() != CV_32FC1 || D.type() != CV_32FC1 || D.total() != 5 || m1type != CV_16SC2 || newK.data != K.data) unsupported("cv::initUndistortRectifyMap");
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      if (R.at<double>(i, j) != (i == j ? 1.0 : 0.0)) unsupported("cv::initUndistortRectifyMap with R != I");
  orc_cam cam{};
  cam.model = 0; cam.width = size.width; cam.height = size.height;
  cam.fx = K.at<float>(0, 0); cam.fy = K.at<float>(1, 1); cam.cx = K.at<float>(0, 2); cam.cy = K.at<float>(1, 2);
  for (int i = 0; i < 5; ++i) cam.d[i] = ((const float*)D.data)[i];
  Mat o1(size.height, size.width, CV_16SC2), o2(size.height, size.width, CV_16UC1);
  orc_init_undistort_maps(&cam, (int16_t*)o1.data, (uint16_t*)o2.data);
  map1 = o1;
  map2 = o2;
}

// camera.cpp:78-81: cv::undistortPoints(src 1x1 CV_32FC2, dst 1x1 CV_32FC2 (user data), cvK_, cvD_)
void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D) {
  if (src.type() != CV_32FC2 || src.total() != 1 || dst.type() != CV_32FC2 || dst.total() != 1 || K.type() != CV_32FC1 || D.type() != CV_32FC1 || D.total() != 5)
    unsupported("cv::undistortPoints");
  const float Kf[4] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2)};
  orc_cv_undistort_point(Kf, (const float*)D.data, ((const float*)src.data)[0], ((const float*)src.data)[1], (float*)dst.data);
}

}  // namespace cv
