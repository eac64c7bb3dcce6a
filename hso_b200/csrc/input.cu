// Row N4 (input side) on sm_100a: what the reference does to a raw image before FrameHandlerMono::addImage (test/test_dataset.cpp:262-283):
//   ImageReader::readImage's cv::resize(image, image, m_img_new_size) (src/ImageReader.cpp:80, INTER_LINEAR)   -> k_resize_u8
//   AbstractCamera::undistortImage = cv::remap(raw, rectified, undist_map1_, undist_map2_, INTER_LINEAR)          -> k_remap_u8
//   (src/camera.cpp:127-131 pinhole, :267-271 FOV, :365-369 equidistant)
// and the CV_16SC2 fixed-point maps the camera constructors build once (host side, below): cv::initUndistortRectifyMap for the pinhole
// model (camera.cpp:47-54), FOVCamera::getRemap / EquidistantCamera::getRemap + cv::convertMaps (camera.cpp:223-265, 317-363).
// Integer pipelines: results are bit-exact against OpenCV (cv2 4.13 golden vectors, tests/golden/cv_golden2.npz).
//
// Both kernels are HBM streaming: per destination pixel k_remap_u8 reads 6 B of map and writes 1 B (the 4 source taps come through
// L1/L2: neighbouring destination pixels hit neighbouring source pixels); each thread produces 4 consecutive pixels from one 16-byte and
// one 8-byte map load and stores one 32-bit word.
#include <cmath>
#include <vector>

#include "hso_internal.h"

namespace hso {

__global__ void __launch_bounds__(128) k_remap_u8(const uint8_t* const* __restrict__ srcs, int sw, int sh, int sstride, const short2* __restrict__ map1,
                                                  const uint16_t* __restrict__ map2, uint8_t* const* __restrict__ dsts, int dw, int dh) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
  if (x0 >= dw) return;
  const uint8_t* __restrict__ src = srcs[blockIdx.z];
  uint8_t* dst = dsts[blockIdx.z] + (size_t)y * dw;
  const size_t o = (size_t)y * dw + x0;
  short2 m[4];
  uint16_t f[4];
  if (x0 + 3 < dw && (o & 3) == 0) {
    const int4 mm = __ldg(reinterpret_cast<const int4*>(map1 + o));
    const uint2 ff = __ldg(reinterpret_cast<const uint2*>(map2 + o));
    m[0] = make_short2((short)(mm.x & 0xFFFF), (short)(mm.x >> 16)); m[1] = make_short2((short)(mm.y & 0xFFFF), (short)(mm.y >> 16));
    m[2] = make_short2((short)(mm.z & 0xFFFF), (short)(mm.z >> 16)); m[3] = make_short2((short)(mm.w & 0xFFFF), (short)(mm.w >> 16));
    f[0] = (uint16_t)(ff.x & 0xFFFF); f[1] = (uint16_t)(ff.x >> 16); f[2] = (uint16_t)(ff.y & 0xFFFF); f[3] = (uint16_t)(ff.y >> 16);
  } else {
    for (int k = 0; k < 4; ++k) {
      const bool in = x0 + k < dw;
      m[k] = in ? map1[o + k] : make_short2(-2, -2);
      f[k] = in ? map2[o + k] : (uint16_t)0;
    }
  }
  uint32_t packed = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int sx = m[k].x, sy = m[k].y;
    const int fxy = f[k] & 1023;
    const int fx = fxy & 31, fy = fxy >> 5;
    // BilinearTab_i (INTER_REMAP_COEF_SCALE = 32768): products of multiples of 1/32, exact
    const int w0 = (32 - fx) * (32 - fy) * 32, w1 = fx * (32 - fy) * 32, w2 = (32 - fx) * fy * 32, w3 = fx * fy * 32;
    int r = 0;
    if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
      const uint8_t* S = src + (size_t)sy * sstride + sx;
      const int val = __ldg(S) * w0 + __ldg(S + 1) * w1 + __ldg(S + sstride) * w2 + __ldg(S + sstride + 1) * w3;
      r = (val + (1 << 14)) >> 15;
    } else if (!(sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0)) {  // BORDER_CONSTANT 0 for the taps outside
      const bool x_0 = sx >= 0 && sx < sw, x_1 = sx + 1 >= 0 && sx + 1 < sw, y_0 = sy >= 0 && sy < sh, y_1 = sy + 1 >= 0 && sy + 1 < sh;
      const int v0 = (x_0 && y_0) ? src[(size_t)sy * sstride + sx] : 0, v1 = (x_1 && y_0) ? src[(size_t)sy * sstride + sx + 1] : 0;
      const int v2 = (x_0 && y_1) ? src[(size_t)(sy + 1) * sstride + sx] : 0, v3 = (x_1 && y_1) ? src[(size_t)(sy + 1) * sstride + sx + 1] : 0;
      r = (v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3 + (1 << 14)) >> 15;
    }
    r = r < 0 ? 0 : (r > 255 ? 255 : r);
    packed |= (uint32_t)r << (8 * k);
  }
  if (x0 + 3 < dw && ((reinterpret_cast<uintptr_t>(dst + x0) & 3) == 0)) {
    *reinterpret_cast<uint32_t*>(dst + x0) = packed;
  } else {
    for (int k = 0; k < 4 && x0 + k < dw; ++k) dst[x0 + k] = (uint8_t)(packed >> (8 * k));
  }
}

// cv::resize(INTER_LINEAR, CV_8UC1): OpenCV's 11-bit fixed-point separable bilinear (same arithmetic as k_resize_level, pyramid.cu).
__global__ void __launch_bounds__(128) k_resize_u8(const uint8_t* const* __restrict__ srcs, int sw, int sh, int sstride, ResizeTabDev tab,
                                                   uint8_t* const* __restrict__ dsts, int dw, int dh) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw || y >= dh) return;
  const uint8_t* __restrict__ src = srcs[blockIdx.z];
  uint8_t* dst = dsts[blockIdx.z];
  if (tab.area_fast) {
    const uint8_t* s = src + (size_t)(2 * y) * sstride + 2 * x;
    dst[(size_t)y * dw + x] = (uint8_t)((s[0] + s[1] + s[sstride] + s[sstride + 1] + 2) >> 2);
    return;
  }
  const int sy0 = min(max(tab.yofs[y], 0), sh - 1), sy1 = min(max(tab.yofs[y] + 1, 0), sh - 1);
  const int sx = tab.xofs[x], sx1 = min(sx + 1, sw - 1);
  const int a0 = tab.ialpha[2 * x], a1 = tab.ialpha[2 * x + 1];
  const int b0 = tab.ibeta[2 * y], b1 = tab.ibeta[2 * y + 1];
  const int r0 = __ldg(src + (size_t)sy0 * sstride + sx) * a0 + __ldg(src + (size_t)sy0 * sstride + sx1) * a1;
  const int r1 = __ldg(src + (size_t)sy1 * sstride + sx) * a0 + __ldg(src + (size_t)sy1 * sstride + sx1) * a1;
  int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
  v = v < 0 ? 0 : (v > 255 ? 255 : v);
  dst[(size_t)y * dw + x] = (uint8_t)v;
}

cudaError_t launch_remap(const uint8_t* const* srcs_dev, int sw, int sh, int sstride, const short2* map1, const uint16_t* map2,
                         uint8_t* const* dsts_dev, int dw, int dh, int B, cudaStream_t stream, uint64_t* launches) {
  dim3 grid((dw + 511) / 512, dh, B);
  k_remap_u8<<<grid, 128, 0, stream>>>(srcs_dev, sw, sh, sstride, map1, map2, dsts_dev, dw, dh);
  ++*launches;
  return cudaGetLastError();
}
cudaError_t launch_resize(const uint8_t* const* srcs_dev, int sw, int sh, int sstride, const ResizeTabDev& tab, uint8_t* const* dsts_dev, int dw,
                          int dh, int B, cudaStream_t stream, uint64_t* launches) {
  dim3 grid((dw + 127) / 128, dh, B);
  k_resize_u8<<<grid, 128, 0, stream>>>(srcs_dev, sw, sh, sstride, tab, dsts_dev, dw, dh);
  ++*launches;
  return cudaGetLastError();
}

// ---- host: the maps the camera constructors build once ------------------------------------------------------------------------------
namespace {
inline int round_half_even(double v) { return (int)std::nearbyint(v); }  // cvRound / saturate_cast<int>(double)
inline short sat16(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }
inline void put_fixed(int iu, int iv, short* m1, uint16_t* m2) {
  m1[0] = sat16(iu >> 5);
  m1[1] = sat16(iv >> 5);
  *m2 = (uint16_t)((iv & 31) * 32 + (iu & 31));
}
}  // namespace

void build_undistort_maps(const hso_cam& cam, std::vector<short>& map1, std::vector<uint16_t>& map2) {
  const int W = cam.width, H = cam.height;
  map1.assign((size_t)2 * W * H, 0);
  map2.assign((size_t)W * H, 0);
  if (cam.model == 0) {
    // cv::initUndistortRectifyMap(cvK_, cvD_, I, cvK_, size, CV_16SC2): K and D are float matrices (src/camera.cpp:43-54)
    const double fx = (double)(float)cam.fx, fy = (double)(float)cam.fy, u0 = (double)(float)cam.cx, v0 = (double)(float)cam.cy;
    const double k1 = (double)(float)cam.d[0], k2 = (double)(float)cam.d[1], p1 = (double)(float)cam.d[2], p2 = (double)(float)cam.d[3],
                 k3 = (double)(float)cam.d[4];
    // inverse of the new camera matrix: cv::invert's closed-form 3x3 branch, cofactor * (1 / det)
    const double a[9] = {fx, 0, u0, 0, fy, v0, 0, 0, 1};
    const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    const double id = 1. / det;
    const double i0 = (a[4] * a[8] - a[5] * a[7]) * id, i1 = (a[2] * a[7] - a[1] * a[8]) * id, i2 = (a[1] * a[5] - a[2] * a[4]) * id;
    const double i3 = (a[5] * a[6] - a[3] * a[8]) * id, i4 = (a[0] * a[8] - a[2] * a[6]) * id, i5 = (a[2] * a[3] - a[0] * a[5]) * id;
    const double i6 = (a[3] * a[7] - a[4] * a[6]) * id, i7 = (a[1] * a[6] - a[0] * a[7]) * id, i8 = (a[0] * a[4] - a[1] * a[3]) * id;
    for (int i = 0; i < H; ++i) {
      double xs = i * i1 + i2, ys = i * i4 + i5, ws = i * i7 + i8;  // running sums along the row, like the published loop
      for (int j = 0; j < W; ++j, xs += i0, ys += i3, ws += i6) {
        const double w = 1. / ws, x = xs * w, y = ys * w;
        const double x2 = x * x, y2 = y * y, r2 = x2 + y2, xy2 = 2 * x * y;
        const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / 1.0;
        const double xd = x * kr + p1 * xy2 + p2 * (r2 + 2 * x2);
        const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * xy2;
        put_fixed(round_half_even((fx * xd + u0) * 32), round_half_even((fy * yd + v0) * 32), &map1[((size_t)i * W + j) * 2], &map2[(size_t)i * W + j]);
      }
    }
    return;
  }
  for (int v = 0; v < H; ++v)
    for (int u = 0; u < W; ++u) {
      float ox, oy;
      const float x = (float)u, y = (float)v;
      const float ix = (float)(((double)x - cam.cx) / cam.fx), iy = (float)(((double)y - cam.cy) / cam.fy);
      if (cam.model == 1) {  // FOVCamera::distortPixelFOV (src/camera.cpp:247-265): float locals against double members
        const float om = (float)cam.d[0];
        const float d2t = (float)(2 * std::tan((double)om / 2));
        const float r = sqrtf(ix * ix + iy * iy);
        const float fac = (r == 0 || om == 0) ? 1 : atanf(r * d2t) / (om * r);
        ox = (float)(cam.fx * (double)fac * (double)ix + cam.cx);
        oy = (float)(cam.fy * (double)fac * (double)iy + cam.cy);
      } else {               // EquidistantCamera::distortPixelEquidistant (src/camera.cpp:342-363)
        const float r = (float)std::sqrt((double)(ix * ix + iy * iy));
        const float th = (float)std::atan((double)r);
        const float th2 = th * th, th4 = th2 * th2, th6 = th4 * th2, th8 = th4 * th4;
        const float thd = (float)((double)th * (1 + cam.d[0] * (double)th2 + cam.d[1] * (double)th4 + cam.d[2] * (double)th6 + cam.d[3] * (double)th8));
        const float sc = (r > 1e-8) ? thd / r : 1.0f;
        ox = (float)(cam.fx * (double)ix * (double)sc + cam.cx);
        oy = (float)(cam.fy * (double)iy * (double)sc + cam.cy);
      }
      // cv::convertMaps(float, float -> CV_16SC2): saturate_cast<int>(x * INTER_TAB_SIZE) evaluated in float
      put_fixed(round_half_even((double)(ox * 32.f)), round_half_even((double)(oy * 32.f)), &map1[((size_t)v * W + u) * 2], &map2[(size_t)v * W + u]);
    }
}

}  // namespace hso
