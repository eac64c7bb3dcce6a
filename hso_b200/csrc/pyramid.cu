// Frame construction on sm_100a: image pyramid + interior statistics (+ optional Sobel images) for a batch of frames.
//
// Replaces frame_utils::createImgPyramid (src/frame.cpp:296-314) -> hso::halfSample (src/vikit/vision.cpp:19-44,70-108) or
// cv::resize(INTER_LINEAR), and Frame::prepareForFeatureDetect (src/frame.cpp:205-246).
//
// k_pyr_tile: one CTA per 128x32 level-0 tile. The tile (+16-byte / 2-row halo for the 5x5 Sobel) is staged into shared
// memory with one TMA bulk copy per row (cp.async.bulk + mbarrier, SASS UBLKCP) when the source rows are 16-byte aligned,
// plain coalesced loads otherwise. From shared memory the CTA (a) writes the level-0 copy with 16-byte stores, (b) runs the
// whole halfSample chain L1..L4 inside the tile (a 16x16 level-0 block is one level-4 pixel; each level is computed from the
// previous level's rounded bytes exactly like the reference chain, with the SSE2 double-rounding or the scalar truncation
// selected per level by `cols % 16`), and (c) reduces sum(I) (exact, integer) and sum(|Sobel5 gradient|) over the 16-px-inset
// interior; the last CTA of a frame folds the per-tile partials in a fixed order (deterministic) into integralImage_/gradMean_.
// HBM traffic per frame: W*H read once, W*H*(1 + 1/4 + 1/16 + 1/64 + 1/256) written.
#include "hso_internal.h"

namespace hso {

constexpr int TW = 128, TH = 32, HX = 16, HY = 2;
constexpr int TPITCH = TW + 2 * HX;  // 160
constexpr int TROWS = TH + 2 * HY;   // 36
constexpr int PYR_THREADS = 256;
static_assert(TW == 128 && TH == 32 && PYR_THREADS == 256, "the tile loops index with shifts for a 128x32 tile");

struct PyrKParams {
  PyrGeom g;
  int src_stride;
  int tiles_x, tiles_y;
  int use_tma;
};

HSO_DEV uint8_t half_px(int t0, int t1, int b0_, int b1_, int sse) {
  if (sse) {
    const int v0 = (t0 + b0_ + 1) >> 1, v1 = (t1 + b1_ + 1) >> 1;  // _mm_avg_epu8 on the row pair
    return (uint8_t)((v0 + v1 + 1) >> 1);                          // _mm_avg_epu16 on the column pair
  }
  return (uint8_t)((t0 + t1 + b0_ + b1_) >> 2);                    // scalar fallback truncates (vision.cpp:100)
}

template <bool HALF>
__global__ void __launch_bounds__(PYR_THREADS) k_pyr_tile(const PyrKParams P, const PyrJobDev* __restrict__ jobs, unsigned int* __restrict__ counters) {
  __shared__ __align__(128) uint8_t tile[TROWS * TPITCH];
  __shared__ __align__(16) int16_t hd[TROWS * TW];
  __shared__ __align__(16) int16_t hs[TROWS * TW];
  __shared__ uint8_t l1[(TH / 2) * (TW / 2)], l2[(TH / 4) * (TW / 4)], l3[(TH / 8) * (TW / 8)];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ double red_g[PYR_THREADS / 32];
  __shared__ unsigned long long red_i[PYR_THREADS / 32];

  const PyrJobDev job = jobs[blockIdx.z];
  const int W = P.g.w[0], H = P.g.h[0];
  const int tx = blockIdx.x, ty = blockIdx.y;
  const int x_base = tx * TW - HX, y_base = ty * TH - HY;
  const int tw = min(TW, W - tx * TW), th = min(TH, H - ty * TH);
  const int tid = threadIdx.x;

  // ---- stage tile + halo ------------------------------------------------------------------------------------------------
  const int xs = max(0, x_base), xe = min(W, x_base + TPITCH);
  const int ys = max(0, y_base), ye = min(H, y_base + TROWS);
  if (P.use_tma) {
    if (tid == 0) { mbar_init(&mbar, 1); mbar_fence_init(); fence_proxy_async(); }
    __syncthreads();
    if (tid == 0) mbar_expect_tx(&mbar, (uint32_t)((xe - xs) * (ye - ys)));
    if (tid < TROWS) {
      const int y = y_base + tid;
      if (y >= ys && y < ye) tma_bulk_g2s(tile + tid * TPITCH + (xs - x_base), job.src + (size_t)y * P.src_stride + xs, (uint32_t)(xe - xs), &mbar);
    }
    // one warp polls the mbarrier; the others wait at the CTA barrier instead of spending issue slots in the spin loop (it was 15 % of the
    // kernel's executed instructions in an issue-bound kernel)
    if (tid < 32) mbar_wait(&mbar, 0);
    __syncthreads();
  } else {
    for (int idx = tid; idx < TROWS * TPITCH; idx += PYR_THREADS) {
      const int r = idx / TPITCH, cix = idx - r * TPITCH;
      const int y = y_base + r, x = x_base + cix;
      if (y >= ys && y < ye && x >= xs && x < xe) tile[idx] = job.src[(size_t)y * P.src_stride + x];
    }
    __syncthreads();
  }

  // ---- level-0 copy -----------------------------------------------------------------------------------------------------
  uint8_t* pyr = job.pyr;
  if (job.src == pyr + P.g.off[0]) {
    // built in place: the upload went straight into the level-0 slot
  } else if ((W & 15) == 0) {
    // (index arithmetic on the full 128x32 tile: shifts instead of divisions by the clipped tile size; edge tiles mask the excess)
    const int chunks = tw >> 4;
    for (int idx = tid; idx < TH * (TW / 16); idx += PYR_THREADS) {
      const int r = idx >> 3, cix = idx & 7;
      if (r >= th || cix >= chunks) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(tile + (r + HY) * TPITCH + HX + cix * 16);
      *reinterpret_cast<uint4*>(pyr + P.g.off[0] + (size_t)(ty * TH + r) * W + tx * TW + cix * 16) = v;
    }
  } else {
    for (int idx = tid; idx < th * tw; idx += PYR_THREADS) {
      const int r = idx / tw, cix = idx - r * tw;
      pyr[P.g.off[0] + (size_t)(ty * TH + r) * W + tx * TW + cix] = tile[(r + HY) * TPITCH + HX + cix];
    }
  }

  // ---- Sobel 5x5 on the tile: horizontal pass (derivative [-1,-2,0,2,1], smoothing [1,4,6,4,1]) ---------------------------
  // Four adjacent pixels per thread from three aligned words of the tile row (bytes x-4 .. x+7); the four leading taps of a pixel are one
  // 4-byte window (PRMT) and one dp4a against the packed weights, the fifth tap (weight 1 in both filters) rides in as the accumulator.
  // A warp covers one tile row per step (32 x 4 pixels), the row index advances by a constant: no per-iteration index arithmetic.
  // (The first version — two pixels per thread from byte loads, idx -> (row, column) per iteration — spent 31 % of the kernel's instructions here.)
  {
    const int g4 = (tid & 31) * 4;
    constexpr int D4 = 0x0200FEFF;  // (-1, -2, 0, 2) as s8x4, little endian
    constexpr int S4 = 0x04060401;  // ( 1,  4, 6, 4)
#pragma unroll
    for (int j = 0; j < (TROWS + 7) / 8; ++j) {
      const int r = (tid >> 5) + 8 * j;
      if (r < TROWS) {
        const uint32_t* wrow = reinterpret_cast<const uint32_t*>(tile + r * TPITCH + HX + g4);
        const uint32_t w0 = wrow[-1], w1 = wrow[0], w2 = wrow[1];
        const uint32_t win0 = __byte_perm(w0, w1, 0x5432), win1 = __byte_perm(w0, w1, 0x6543), win3 = __byte_perm(w1, w2, 0x4321);
        const int t0 = (int)__byte_perm(w1, 0, 0x4442), t1 = (int)__byte_perm(w1, 0, 0x4443), t2 = (int)__byte_perm(w2, 0, 0x4440),
                  t3 = (int)__byte_perm(w2, 0, 0x4441);
        int d0, d1, d2, d3, s0, s1, s2, s3;
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d0) : "r"(win0), "r"(D4), "r"(t0));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d1) : "r"(win1), "r"(D4), "r"(t1));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d2) : "r"(w1), "r"(D4), "r"(t2));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d3) : "r"(win3), "r"(D4), "r"(t3));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(s0) : "r"(win0), "r"(S4), "r"(t0));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(s1) : "r"(win1), "r"(S4), "r"(t1));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(s2) : "r"(w1), "r"(S4), "r"(t2));
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(s3) : "r"(win3), "r"(S4), "r"(t3));
        // |d| <= 6 * 255, s <= 16 * 255: int16; four results as one 8-byte store
        *reinterpret_cast<uint2*>(hd + r * TW + g4) = make_uint2(__byte_perm((uint32_t)d0, (uint32_t)d1, 0x5410), __byte_perm((uint32_t)d2, (uint32_t)d3, 0x5410));
        *reinterpret_cast<uint2*>(hs + r * TW + g4) = make_uint2(__byte_perm((uint32_t)s0, (uint32_t)s1, 0x5410), __byte_perm((uint32_t)s2, (uint32_t)s3, 0x5410));
      }
    }
  }

  // ---- halfSample chain inside the tile ---------------------------------------------------------------------------------
  if (HALF) {
    // L1: (th/2) x (tw/2), 4 px per thread-iteration
    {
      const int w1 = P.g.w[1], sse = P.g.sse_rounding[1];
      const int cw = tw >> 3;  // groups of 4 output px per row
      for (int idx = tid; idx < (TH / 2) * (TW / 8); idx += PYR_THREADS) {
        const int r = idx >> 4, cg4 = idx & 15;
        if (r >= (th >> 1) || cg4 >= cw) continue;
        const uint8_t* t = tile + (2 * r + HY) * TPITCH + HX + cg4 * 8;
        const uint8_t* b = t + TPITCH;
        uchar4 o;
        o.x = half_px(t[0], t[1], b[0], b[1], sse);
        o.y = half_px(t[2], t[3], b[2], b[3], sse);
        o.z = half_px(t[4], t[5], b[4], b[5], sse);
        o.w = half_px(t[6], t[7], b[6], b[7], sse);
        *reinterpret_cast<uchar4*>(l1 + r * (TW / 2) + cg4 * 4) = o;
        *reinterpret_cast<uchar4*>(pyr + P.g.off[1] + (size_t)(ty * (TH / 2) + r) * w1 + tx * (TW / 2) + cg4 * 4) = o;
      }
    }
    __syncthreads();
    if (P.g.n_levels > 2) {
      const int w2 = P.g.w[2], sse = P.g.sse_rounding[2];
      const int cw = tw >> 2;
      for (int idx = tid; idx < (TH / 4) * (TW / 4); idx += PYR_THREADS) {
        const int r = idx >> 5, cix = idx & 31;
        if (r >= (th >> 2) || cix >= cw) continue;
        const uint8_t* t = l1 + (2 * r) * (TW / 2) + 2 * cix;
        const uint8_t* b = t + TW / 2;
        const uint8_t o = half_px(t[0], t[1], b[0], b[1], sse);
        l2[r * (TW / 4) + cix] = o;
        pyr[P.g.off[2] + (size_t)(ty * (TH / 4) + r) * w2 + tx * (TW / 4) + cix] = o;
      }
    }
    __syncthreads();
    if (P.g.n_levels > 3) {
      const int w3 = P.g.w[3], sse = P.g.sse_rounding[3];
      const int cw = tw >> 3;
      for (int idx = tid; idx < (TH / 8) * (TW / 8); idx += PYR_THREADS) {
        const int r = idx >> 4, cix = idx & 15;
        if (r >= (th >> 3) || cix >= cw) continue;
        const uint8_t* t = l2 + (2 * r) * (TW / 4) + 2 * cix;
        const uint8_t* b = t + TW / 4;
        const uint8_t o = half_px(t[0], t[1], b[0], b[1], sse);
        l3[r * (TW / 8) + cix] = o;
        pyr[P.g.off[3] + (size_t)(ty * (TH / 8) + r) * w3 + tx * (TW / 8) + cix] = o;
      }
    }
    __syncthreads();
    if (P.g.n_levels > 4) {
      const int w4 = P.g.w[4], sse = P.g.sse_rounding[4];
      const int cw = tw >> 4;
      for (int idx = tid; idx < (TH / 16) * (TW / 16); idx += PYR_THREADS) {
        const int r = idx >> 3, cix = idx & 7;
        if (r >= (th >> 4) || cix >= cw) continue;
        const uint8_t* t = l3 + (2 * r) * (TW / 8) + 2 * cix;
        const uint8_t* b = t + TW / 8;
        pyr[P.g.off[4] + (size_t)(ty * (TH / 16) + r) * w4 + tx * (TW / 16) + cix] = half_px(t[0], t[1], b[0], b[1], sse);
      }
    }
  } else {
    __syncthreads();
  }

  // ---- vertical Sobel pass + interior sums (src/frame.cpp:223-245) ---------------------------------------------------------
  // Each thread walks one column of one 16-row half of the tile and keeps the last five horizontal results in registers: two shared-memory
  // loads per pixel instead of nine.
  float gsum = 0.f;
  unsigned int isum = 0;
  {
    const int col = tid & (TW - 1), r0 = (tid >> 7) * (TH / 2);
    const int x = tx * TW + col;
    if (x >= 16 && x < W - 16) {  // also masks the excess of edge tiles (x >= W)
      const int16_t* dc = hd + r0 * TW + col;  // tile row r is image row y-2
      const int16_t* sc = hs + r0 * TW + col;
      const uint8_t* tc = tile + (r0 + HY) * TPITCH + HX + col;
      int d0 = dc[0], d1 = dc[TW], d2 = dc[2 * TW], d3 = dc[3 * TW];
      int s0 = sc[0], s1 = sc[TW], s2 = sc[2 * TW], s3 = sc[3 * TW];
#pragma unroll
      for (int k = 0; k < TH / 2; ++k) {
        const int d4 = dc[(k + 4) * TW], s4 = sc[(k + 4) * TW];
        const int y = ty * TH + r0 + k;
        {
          // interior rows only (16 <= y < H - 16), branch-free: the row is evaluated anyway and masked out of the two sums
          const bool in = (unsigned)(y - 16) < (unsigned)(H - 32);
          const int gx = (d0 + d4) + 4 * (d1 + d3) + 6 * d2;
          const int gy = (s4 - s0) + 2 * (s3 - s1);
          const float fx = (float)gx, fy = (float)gy;
          // |grad| through the single-instruction square root (<= 2 ulp): it feeds a mean over 2.7e5 pixels that is compared at 2.5e-4
          float mag;
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(fx * fx + fy * fy));
          gsum += in ? mag : 0.f;
          isum += in ? (unsigned)tc[k * TPITCH] : 0u;
        }
        d0 = d1; d1 = d2; d2 = d3; d3 = d4;
        s0 = s1; s1 = s2; s2 = s3; s3 = s4;
      }
    }
  }
  {
    double g = (double)gsum;
    unsigned long long iv = isum;
    for (int o = 16; o > 0; o >>= 1) {
      g += __shfl_xor_sync(0xffffffffu, g, o);
      iv += __shfl_xor_sync(0xffffffffu, iv, o);
    }
    if ((tid & 31) == 0) { red_g[tid >> 5] = g; red_i[tid >> 5] = iv; }
  }
  __syncthreads();
  const int n_tiles = P.tiles_x * P.tiles_y;
  const int tile_id = ty * P.tiles_x + tx;
  if (tid < 32) {  // warp 0 only: the other warps leave without waiting for the counter's round trip
    unsigned int prev = 0;
    if (tid == 0) {
      double g = 0;
      unsigned long long iv = 0;
      for (int w = 0; w < PYR_THREADS / 32; ++w) { g += red_g[w]; iv += red_i[w]; }
      job.sums[2 * tile_id] = g;
      job.sums[2 * tile_id + 1] = (double)iv;
      __threadfence();
      prev = atomicAdd(&counters[blockIdx.z], 1u);
    }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev != (unsigned)(n_tiles - 1)) return;
    // last CTA of the frame: fold the per-tile partials, lane-strided loads + a fixed-order butterfly (deterministic)
    __threadfence();
    double gt = 0, it = 0;
    const volatile double* sums = job.sums;
    for (int t = tid; t < n_tiles; t += 32) { gt += sums[2 * t]; it += sums[2 * t + 1]; }
    for (int o = 16; o > 0; o >>= 1) {
      gt += __shfl_xor_sync(0xffffffffu, gt, o);
      it += __shfl_xor_sync(0xffffffffu, it, o);
    }
    if (tid == 0) {
      const int cnt = (W - 32) * (H - 32);
      float integral = (float)it / (float)cnt;
      float gm = (float)gt / (float)cnt;
      gm /= 30.f;
      if (gm > 20.f) gm = 20.f;
      if (gm < 7.f) gm = 7.f;
      job.stats[0] = integral;
      job.stats[1] = gm;
      counters[blockIdx.z] = 0;  // ready for the next launch
    }
  }
}

// cv::resize(INTER_LINEAR, CV_8UC1) for one level of the non-%16 pyramid path (src/frame.cpp:309-311): OpenCV's 11-bit
// fixed-point separable bilinear, or the exact-2x decimation fast path (a+b+c+d+2)>>2. Coefficient tables come from the host.
__global__ void k_resize_level(const PyrJobDev* __restrict__ jobs, size_t src_off, int sw, int sh, size_t dst_off, int dw, int dh, ResizeTabDev tab) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw || y >= dh) return;
  const uint8_t* src = jobs[blockIdx.z].pyr + src_off;
  uint8_t* dst = jobs[blockIdx.z].pyr + dst_off;
  if (tab.area_fast) {
    const uint8_t* s = src + (size_t)(2 * y) * sw + 2 * x;
    dst[(size_t)y * dw + x] = (uint8_t)((s[0] + s[1] + s[sw] + s[sw + 1] + 2) >> 2);
    return;
  }
  const int sy0 = min(max(tab.yofs[y], 0), sh - 1), sy1 = min(max(tab.yofs[y] + 1, 0), sh - 1);
  const int sx = tab.xofs[x], sx1 = min(sx + 1, sw - 1);
  const int a0 = tab.ialpha[2 * x], a1 = tab.ialpha[2 * x + 1];
  const int b0_ = tab.ibeta[2 * y], b1_ = tab.ibeta[2 * y + 1];
  const int r0 = src[(size_t)sy0 * sw + sx] * a0 + src[(size_t)sy0 * sw + sx1] * a1;
  const int r1 = src[(size_t)sy1 * sw + sx] * a0 + src[(size_t)sy1 * sw + sx1] * a1;
  int v = (((b0_ * (r0 >> 4)) >> 16) + ((b1_ * (r1 >> 4)) >> 16) + 2) >> 2;
  v = v < 0 ? 0 : (v > 255 ? 255 : v);
  dst[(size_t)y * dw + x] = (uint8_t)v;
}

// cv::Sobel(CV_16S, ksize 5, BORDER_REPLICATE) images for levels 0..2 (src/frame.cpp:216-220), only when materialised.
__global__ void k_sobel5(const PyrJobDev* __restrict__ jobs, size_t lvl_off, int w, int h, size_t sob_off) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const uint8_t* img = jobs[blockIdx.z].pyr + lvl_off;
  int16_t* gxo = jobs[blockIdx.z].sobel + sob_off;
  int16_t* gyo = gxo + (size_t)w * h;
  const int D[5] = {-1, -2, 0, 2, 1}, S[5] = {1, 4, 6, 4, 1};
  int gx = 0, gy = 0;
#pragma unroll
  for (int j = -2; j <= 2; ++j) {
    const int yy = min(max(y + j, 0), h - 1);
    int sd = 0, ss = 0;
#pragma unroll
    for (int k = -2; k <= 2; ++k) {
      const int xx = min(max(x + k, 0), w - 1);
      const int v = img[(size_t)yy * w + xx];
      sd += D[k + 2] * v;
      ss += S[k + 2] * v;
    }
    gx += S[j + 2] * sd;
    gy += D[j + 2] * ss;
  }
  gxo[(size_t)y * w + x] = (int16_t)max(-32768, min(32767, gx));
  gyo[(size_t)y * w + x] = (int16_t)max(-32768, min(32767, gy));
}

int pyramid_tiles(const PyrGeom& g) { return ((g.w[0] + TW - 1) / TW) * ((g.h[0] + TH - 1) / TH); }

cudaError_t launch_pyramid(const PyrGeom& g, const PyrJobDev* jobs_dev, int B, int src_stride, const ResizeTabDev* tabs, int store_sobel,
                           unsigned int* counters_dev, int src_aligned16, cudaStream_t stream, uint64_t* launches) {
  PyrKParams P;
  P.g = g;
  P.src_stride = src_stride;
  P.tiles_x = (g.w[0] + TW - 1) / TW;
  P.tiles_y = (g.h[0] + TH - 1) / TH;
  P.use_tma = (src_aligned16 && (src_stride % 16) == 0 && (g.w[0] % 16) == 0) ? 1 : 0;
  dim3 grid(P.tiles_x, P.tiles_y, B);
  if (g.half_path) k_pyr_tile<true><<<grid, PYR_THREADS, 0, stream>>>(P, jobs_dev, counters_dev);
  else k_pyr_tile<false><<<grid, PYR_THREADS, 0, stream>>>(P, jobs_dev, counters_dev);
  ++*launches;
  if (!g.half_path) {
    for (int l = 1; l < g.n_levels; ++l) {
      dim3 gr((g.w[l] + 127) / 128, g.h[l], B);
      k_resize_level<<<gr, 128, 0, stream>>>(jobs_dev, g.off[l - 1], g.w[l - 1], g.h[l - 1], g.off[l], g.w[l], g.h[l], tabs[l]);
      ++*launches;
    }
  }
  if (store_sobel) {
    size_t so = 0;
    for (int l = 0; l < 3 && l < g.n_levels; ++l) {
      dim3 gr((g.w[l] + 127) / 128, g.h[l], B);
      k_sobel5<<<gr, 128, 0, stream>>>(jobs_dev, g.off[l], g.w[l], g.h[l], so);
      ++*launches;
      so += (size_t)2 * g.w[l] * g.h[l];
    }
  }
  return cudaGetLastError();
}

}  // namespace hso
