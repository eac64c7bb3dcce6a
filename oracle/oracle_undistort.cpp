// ORACLE — test infrastructure, never linked into or called by the product (hso_b200/).
// CPU restatement of row N4 (input side): the undistortion maps the reference's camera models build at construction
// (src/camera.cpp:47-54 PinholeCamera: cv::initUndistortRectifyMap(cvK_, cvD_, I, cvK_, size, CV_16SC2); :223-245 FOVCamera::getRemap +
// distortPixelFOV :247-265; :317-340 EquidistantCamera::getRemap + distortPixelEquidistant :342-363, both through cv::convertMaps) and
// cv::remap(raw, rectified, map1, map2, INTER_LINEAR) (:127-131, :267-271, :365-369).
//
// OpenCV is not vendored in /root/reference (unpinned, README "3.2.0"): initUndistortRectifyMap, convertMaps and the fixed-point bilinear
// remap are restated from the published algorithms (modules/calib3d/src/undistort.dispatch.cpp, modules/imgproc/src/imgwarp.cpp:
// INTER_BITS = 5, INTER_TAB_SIZE = 32, INTER_REMAP_COEF_BITS = 15, BORDER_CONSTANT 0) and pinned against cv2 4.13 golden vectors
// (tests/golden/cv_golden2.npz).
#include <cmath>
#include <cstring>

#include "hso_oracle.h"

namespace {

inline int cv_round(double v) { return (int)std::nearbyint(v); }  // cvRound: round half to even (SSE2 cvtsd2si)
inline int sat_int(double v) { return cv_round(v); }               // saturate_cast<int>(double)

inline void store_fixed(double u, double v, int16_t* m1, uint16_t* m2) {
  const int iu = sat_int(u * 32), iv = sat_int(v * 32);
  m1[0] = (int16_t)(iu >> 5);
  m1[1] = (int16_t)(iv >> 5);
  *m2 = (uint16_t)((iv & 31) * 32 + (iu & 31));
}

}  // namespace

extern "C" {

// cv::convertMaps(CV_32FC1, CV_32FC1 -> CV_16SC2, CV_16UC1): ix = saturate_cast<int>(x * INTER_TAB_SIZE) in float, then the short cast
void orc_convert_maps(const float* mapx, const float* mapy, int n, int16_t* map1, uint16_t* map2) {
  for (int i = 0; i < n; ++i) {
    const int iu = cv_round((double)(mapx[i] * 32.f)), iv = cv_round((double)(mapy[i] * 32.f));
    const int su = iu >> 5, sv = iv >> 5;
    map1[2 * i] = (int16_t)(su < -32768 ? -32768 : (su > 32767 ? 32767 : su));      // saturate_cast<short>
    map1[2 * i + 1] = (int16_t)(sv < -32768 ? -32768 : (sv > 32767 ? 32767 : sv));
    map2[i] = (uint16_t)((iv & 31) * 32 + (iu & 31));
  }
}

// map1: [h][w][2] int16 (integer source pixel), map2: [h][w] uint16 (5+5 fractional bits). Returns 0, or 1 if the model has no map.
int orc_init_undistort_maps(const orc_cam* cam, int16_t* map1, uint16_t* map2) {
  const int W = cam->width, H = cam->height;
  if (cam->model == 0) {
    // cv::initUndistortRectifyMap with K, D as float matrices (camera.cpp:43-45), R = I, newCameraMatrix = K
    const double fx = (double)(float)cam->fx, fy = (double)(float)cam->fy, u0 = (double)(float)cam->cx, v0 = (double)(float)cam->cy;
    const double k1 = (double)(float)cam->d[0], k2 = (double)(float)cam->d[1], p1 = (double)(float)cam->d[2], p2 = (double)(float)cam->d[3],
                 k3 = (double)(float)cam->d[4];
    // iR = (Ar * R).inv(DECOMP_LU): cv::invert's closed-form 3x3 branch (cofactors times 1/det)
    const double m[9] = {fx, 0, u0, 0, fy, v0, 0, 0, 1};
    const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    const double d = 1. / det;
    double ir[9];
    ir[0] = (m[4] * m[8] - m[5] * m[7]) * d; ir[1] = (m[2] * m[7] - m[1] * m[8]) * d; ir[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    ir[3] = (m[5] * m[6] - m[3] * m[8]) * d; ir[4] = (m[0] * m[8] - m[2] * m[6]) * d; ir[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    ir[6] = (m[3] * m[7] - m[4] * m[6]) * d; ir[7] = (m[1] * m[6] - m[0] * m[7]) * d; ir[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    for (int i = 0; i < H; ++i) {
      double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
      for (int j = 0; j < W; ++j, _x += ir[0], _y += ir[3], _w += ir[6]) {
        const double w = 1. / _w, x = _x * w, y = _y * w;
        const double x2 = x * x, y2 = y * y;
        const double r2 = x2 + y2, _2xy = 2 * x * y;
        const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((0 * r2 + 0) * r2 + 0) * r2);
        const double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2));
        const double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy);
        const double u = fx * xd + u0;
        const double v = fy * yd + v0;
        store_fixed(u, v, map1 + ((size_t)i * W + j) * 2, map2 + (size_t)i * W + j);
      }
    }
    return 0;
  }
  if (cam->model == 1 || cam->model == 2) {
    for (int v = 0; v < H; ++v)
      for (int u = 0; u < W; ++u) {
        float ox, oy;
        if (cam->model == 1) {
          // FOVCamera::distortPixelFOV — camera.cpp:247-265 (float locals, double members)
          const float dist = (float)cam->d[0];
          const float d2t = (float)(2 * std::tan((double)dist / 2));
          const float x = (float)u, y = (float)v;
          float ix = (float)(((double)x - cam->cx) / cam->fx);
          float iy = (float)(((double)y - cam->cy) / cam->fy);
          const float r = sqrtf(ix * ix + iy * iy);
          const float fac = (r == 0 || dist == 0) ? 1 : atanf(r * d2t) / (dist * r);
          ix = (float)(cam->fx * (double)fac * (double)ix + cam->cx);
          iy = (float)(cam->fy * (double)fac * (double)iy + cam->cy);
          ox = ix; oy = iy;
        } else {
          // EquidistantCamera::distortPixelEquidistant — camera.cpp:342-363
          const float x = (float)u, y = (float)v;
          const float ix = (float)(((double)x - cam->cx) / cam->fx);
          const float iy = (float)(((double)y - cam->cy) / cam->fy);
          const float r = (float)std::sqrt((double)(ix * ix + iy * iy));
          const float theta = (float)std::atan((double)r);
          const float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
          const float thetad = (float)((double)theta * (1 + cam->d[0] * (double)theta2 + cam->d[1] * (double)theta4 + cam->d[2] * (double)theta6 +
                                                        cam->d[3] * (double)theta8));
          const float scaling = (r > 1e-8) ? thetad / r : 1.0f;
          ox = (float)(cam->fx * (double)ix * (double)scaling + cam->cx);
          oy = (float)(cam->fy * (double)iy * (double)scaling + cam->cy);
        }
        orc_convert_maps(&ox, &oy, 1, map1 + ((size_t)v * W + u) * 2, map2 + (size_t)v * W + u);
      }
    return 0;
  }
  return 1;
}

// cv::remap(src, dst, map1 (CV_16SC2), map2 (CV_16UC1), INTER_LINEAR, BORDER_CONSTANT, 0) for CV_8UC1 — imgwarp.cpp remapBilinear
void orc_remap_linear_u8(const uint8_t* src, int sw, int sh, int sstride, const int16_t* map1, const uint16_t* map2, int dw, int dh, uint8_t* dst) {
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      const int sx = map1[((size_t)y * dw + x) * 2], sy = map1[((size_t)y * dw + x) * 2 + 1];
      const int fxy = map2[(size_t)y * dw + x] & 1023;
      const int fx = fxy & 31, fy = fxy >> 5;
      // BilinearTab_i: saturate_cast<short>(w * 32768), w products of multiples of 1/32 — exact, the table sums to 32768 without correction
      const int w0 = (32 - fx) * (32 - fy) * 32, w1 = fx * (32 - fy) * 32, w2 = (32 - fx) * fy * 32, w3 = fx * fy * 32;
      int val;
      if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
        const uint8_t* S = src + (size_t)sy * sstride + sx;
        val = S[0] * w0 + S[1] * w1 + S[sstride] * w2 + S[sstride + 1] * w3;
      } else if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) {
        dst[(size_t)y * dw + x] = 0;
        continue;
      } else {
        auto at = [&](int xx, int yy) -> int { return (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sstride + xx] : 0; };
        val = at(sx, sy) * w0 + at(sx + 1, sy) * w1 + at(sx, sy + 1) * w2 + at(sx + 1, sy + 1) * w3;
      }
      const int r = (val + (1 << 14)) >> 15;  // FixedPtCast<int, uchar, INTER_REMAP_COEF_BITS>
      dst[(size_t)y * dw + x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

}  // extern "C"
