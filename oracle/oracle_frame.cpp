// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of Frame construction (a1, a2) and camera/SE3 helpers.
// Follows: src/frame.cpp:205-246,296-314 ; src/vikit/vision.cpp:19-44,70-108 ; src/camera.cpp:94-125,199-221,307-315.
// Third-party arithmetic restated from published algorithms (OpenCV is NOT vendored in /root/reference and is
// unpinned there — CMakeLists.txt:44-50, README "tested 3.2.0"): cv::Sobel(ksize=5) and cv::resize(INTER_LINEAR, 8UC1).
// Both are pinned against golden vectors produced by cv2 4.13 (tests/golden/make_cv_golden.py).
#include "hso_oracle.h"
#include "oracle_math.hpp"

using namespace orc;

extern "C" {

void orc_se3_exp(const double tangent[6], double rt_out[12]) { SE3::exp(tangent).to_rt(rt_out); }
void orc_se3_log(const double rt[12], double tangent_out[6]) { SE3::from_rt(rt).log(tangent_out); }
void orc_se3_mul(const double a[12], const double b[12], double out[12]) { SE3::from_rt(a).mul(SE3::from_rt(b)).to_rt(out); }
void orc_se3_inverse(const double a[12], double out[12]) { SE3::from_rt(a).inverse().to_rt(out); }
void orc_se3_from_qt(const double q[4], const double t[3], double rt_out[12]) {
  SE3 s;
  s.q = Quat{q[0], q[1], q[2], q[3]};
  s.q.normalize();
  s.t = {t[0], t[1], t[2]};
  s.to_rt(rt_out);
}
void orc_ldlt_solve7(const double* A, const double* b, double* x) { ldlt_solve<7>(A, b, x); }
void orc_ldlt_solve6(const double* A, const double* b, double* x) { ldlt_solve<6>(A, b, x); }

// AbstractCamera::world2cam(Vector3d): src/camera.cpp:94-125 (pinhole, radtan inside), :199-221 (FOV), :307-315 (equidistant).
void orc_world2cam(const orc_cam* cam, const double xyz[3], double px[2]) {
  const double u = xyz[0] / xyz[2], v = xyz[1] / xyz[2];
  if (cam->model == 0) {
    const bool distortion = std::fabs(cam->d[0]) > 0.0000001;  // camera.cpp:36
    if (!distortion) {
      px[0] = cam->fx * u + cam->cx;
      px[1] = cam->fy * v + cam->cy;
    } else {
      const double* d = cam->d;
      double x = u, y = v;
      double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
      double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
      double cdist = 1 + d[0] * r2 + d[1] * r4 + d[4] * r6;
      double xd = x * cdist + d[2] * a1 + d[3] * a2;
      double yd = y * cdist + d[2] * a3 + d[3] * a1;
      px[0] = xd * cam->fx + cam->cx;
      px[1] = yd * cam->fy + cam->cy;
    }
  } else if (cam->model == 1) {
    if (cam->undistort) {
      px[0] = cam->fx * u + cam->cx;
      px[1] = cam->fy * v + cam->cy;
    } else {
      const double omega = cam->d[0];
      double dist = std::sqrt(u * u + v * v);
      double ratio = (omega == 0 || dist == 0) ? 1 : std::atan(2 * dist * std::tan(omega / 2)) / (dist * omega);
      px[0] = ratio * cam->fx * u + cam->cx;
      px[1] = ratio * cam->fy * v + cam->cy;
    }
  } else {
    px[0] = cam->fx * u + cam->cx;
    px[1] = cam->fy * v + cam->cy;
  }
}

// hso::halfSample — src/vikit/vision.cpp:70-108. On x86 the SSE2 kernel (vision.cpp:19-44) is taken iff the buffers are
// 16-byte aligned (cv::Mat allocations always are) and in.cols % 16 == 0: out = avg16(avg8(top,bottom) even, odd) where
// avg rounds half up; otherwise the scalar loop truncates (a+b+c+d)/4.
void orc_half_sample(const uint8_t* in, int w, int h, uint8_t* out, int mode) {
  const int ow = w / 2, oh = h / 2;
  const bool sse = (mode == 1) || (mode == -1 && (w % 16) == 0);
  if (sse) {
    // halfSampleSSE2 walks sw = w>>4 blocks of 16 input pixels per row pair.
    for (int y = 0; y < (h >> 1); ++y) {
      const uint8_t* top = in + (size_t)(2 * y) * w;
      const uint8_t* bot = top + w;
      uint8_t* o = out + (size_t)y * ow;
      for (int x = 0; x < ((w >> 4) << 3); ++x) {
        unsigned v0 = (top[2 * x] + bot[2 * x] + 1u) >> 1;          // _mm_avg_epu8
        unsigned v1 = (top[2 * x + 1] + bot[2 * x + 1] + 1u) >> 1;
        o[x] = (uint8_t)((v0 + v1 + 1u) >> 1);                       // _mm_avg_epu16
      }
    }
  } else {
    for (int y = 0; y < oh; ++y) {
      const uint8_t* top = in + (size_t)(2 * y) * w;
      const uint8_t* bot = top + w;
      uint8_t* o = out + (size_t)y * ow;
      for (int x = 0; x < ow; ++x)
        o[x] = (uint8_t)(((uint16_t)top[2 * x] + top[2 * x + 1] + bot[2 * x] + bot[2 * x + 1]) / 4);
    }
  }
}

static inline int cv_round_f(double v) { return (int)std::nearbyint(v); }  // cvRound: round-half-even (default FP mode)
static inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

// cv::resize(..., INTER_LINEAR) for CV_8UC1. Published algorithm (opencv/modules/imgproc/src/resize.cpp):
//  * exact 2x decimation is rerouted to INTER_AREA's fast path: (a+b+c+d+2)>>2;
//  * otherwise 11-bit fixed-point separable bilinear: horizontal pass into int rows with short coefficients
//    (cvRound(w*2048)), vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  const double inv_fx = (double)sw / dw, inv_fy = (double)sh / dh;  // scale_x, scale_y
  const int iscale_x = (int)std::lrint(inv_fx) /*saturate_cast<int>*/, iscale_y = (int)std::lrint(inv_fy);
  const bool is_area_fast = std::fabs(inv_fx - iscale_x) < 2.220446049250313e-16 && std::fabs(inv_fy - iscale_y) < 2.220446049250313e-16;
  if (is_area_fast && iscale_x == 2 && iscale_y == 2) {
    for (int y = 0; y < dh; ++y)
      for (int x = 0; x < dw; ++x) {
        const uint8_t* s = src + (size_t)(2 * y) * sw + 2 * x;
        dst[(size_t)y * dw + x] = (uint8_t)((s[0] + s[1] + s[sw] + s[sw + 1] + 2) >> 2);
      }
    return;
  }
  const int ONE = 2048;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> ialpha(2 * dw), ibeta(2 * dh);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * inv_fx - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    ialpha[2 * dx] = sat_short(cv_round_f((1.f - fx) * ONE));
    ialpha[2 * dx + 1] = sat_short(cv_round_f(fx * ONE));
  }
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * inv_fy - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    ibeta[2 * dy] = sat_short(cv_round_f((1.f - fy) * ONE));
    ibeta[2 * dy + 1] = sat_short(cv_round_f(fy * ONE));
  }
  std::vector<int> r0(dw), r1(dw);
  for (int dy = 0; dy < dh; ++dy) {
    int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
    int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
    const uint8_t* S0 = src + (size_t)sy0 * sw;
    const uint8_t* S1 = src + (size_t)sy1 * sw;
    for (int dx = 0; dx < dw; ++dx) {
      int sx = xofs[dx];
      int sx1 = std::min(sx + 1, sw - 1);
      int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
      r0[dx] = S0[sx] * a0 + S0[sx1] * a1;
      r1[dx] = S1[sx] * a0 + S1[sx1] * a1;
    }
    int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
    for (int dx = 0; dx < dw; ++dx) {
      int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
      dst[(size_t)dy * dw + dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
}

// frame_utils::createImgPyramid — src/frame.cpp:296-314.
int orc_create_pyramid(const uint8_t* img, int W, int H, int n_levels, uint8_t* out, int* lw, int* lh) {
  lw[0] = W; lh[0] = H;
  const bool half_path = (W % 16 == 0) && (H % 16 == 0);
  const uint8_t* prev = img;
  uint8_t* o = out;
  for (int i = 1; i < n_levels; ++i) {
    if (half_path) {
      lw[i] = lw[i - 1] / 2; lh[i] = lh[i - 1] / 2;
      orc_half_sample(prev, lw[i - 1], lh[i - 1], o, -1);
    } else {
      float scale = 1.0 / (1 << i);
      lw[i] = cv_round_f((float)W * scale);
      lh[i] = cv_round_f((float)H * scale);
      orc_resize_linear_u8(prev, lw[i - 1], lh[i - 1], o, lw[i], lh[i]);
    }
    prev = o;
    o += (size_t)lw[i] * lh[i];
  }
  return half_path ? 0 : 1;
}

// cv::Sobel(src, dst, CV_16S, dx, dy, ksize=5, scale=1, delta=0, BORDER_REPLICATE) — src/frame.cpp:216-220.
// Separable kernels from getSobelKernels: derivative [-1,-2,0,2,1], smoothing [1,4,6,4,1].
void orc_sobel5(const uint8_t* img, int w, int h, int16_t* gx, int16_t* gy) {
  static const int D[5] = {-1, -2, 0, 2, 1};
  static const int S[5] = {1, 4, 6, 4, 1};
  std::vector<int> hx((size_t)w * h), hs((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* r = img + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      int sd = 0, ss = 0;
      for (int k = -2; k <= 2; ++k) {
        int xx = std::min(std::max(x + k, 0), w - 1);
        sd += D[k + 2] * r[xx];
        ss += S[k + 2] * r[xx];
      }
      hx[(size_t)y * w + x] = sd;
      hs[(size_t)y * w + x] = ss;
    }
  }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int sx = 0, sy = 0;
      for (int k = -2; k <= 2; ++k) {
        int yy = std::min(std::max(y + k, 0), h - 1);
        sx += S[k + 2] * hx[(size_t)yy * w + x];
        sy += D[k + 2] * hs[(size_t)yy * w + x];
      }
      gx[(size_t)y * w + x] = sat_short(sx);
      gy[(size_t)y * w + x] = sat_short(sy);
    }
}

// Frame::prepareForFeatureDetect statistics — src/frame.cpp:223-245: float running sums in raster order over the
// 16-px-inset interior of level 0.
void orc_frame_stats(const uint8_t* img, const int16_t* gx, const int16_t* gy, int w, int h, float* integral, float* grad_mean) {
  float intSum = 0, gradSum = 0;
  int sum = 0;
  for (int y = 16; y < h - 16; y++)
    for (int x = 16; x < w - 16; x++) {
      sum++;
      float gradx = gx[(size_t)y * w + x];
      float grady = gy[(size_t)y * w + x];
      gradSum += sqrtf(gradx * gradx + grady * grady);
      intSum += img[(size_t)y * w + x];
    }
  *integral = intSum / sum;
  float gm = gradSum / sum;
  gm /= 30;
  if (gm > 20) gm = 20;
  if (gm < 7) gm = 7;
  *grad_mean = gm;
}

}  // extern "C"
