// FAST-9 corner detection on sm_100a over a device-resident pyramid level — row N2 of the scope table: the detector part of
// FeatureExtractor::fastDetectST (src/feature_detection.cpp:498-523): fast_corner_detect_9_sse2 (thirdparty/fast/src/faster_corner_9_sse.cpp),
// fast_corner_score_9 (thirdparty/fast/src/fast_9_score.cpp), fast_nonmax_3x3 (thirdparty/fast/src/nonmax_3x3.cpp), the 8-px border filter
// and hso::shiTomasiScore (src/vikit/vision.cpp:111-151). Integer work, bit-exact (coordinates, scores, survivors, order).
//
//   k_fast_score : one thread per pixel, 32x8 tile + 3-px halo in shared memory. The 16 circle pixels give a bright and a dark 16-bit mask;
//                  "9 contiguous" is a doubled-mask AND-shift test; the score (largest barrier at which the pixel is still a corner — what
//                  the generated decision tree climbs to) is a binary search with the same predicate. Writes an int16 score map (-1 = no corner).
//   k_fast_rows  : one CTA per row: 3x3 non-max (a neighbour corner with score >= kills), border filter, ordered compaction of the row.
//   k_fast_gather: one CTA per row: row offset = sum of the counts above (raster order like the reference's vectors), Shi-Tomasi score, output.
// HBM bytes per level: w*h read + 2*w*h score map written/read (+ the few survivors) — a streaming kernel; measured in bench.py other_rows.
#include "hso_internal.h"

namespace hso {

constexpr int FT_X = 32, FT_Y = 8, FT_H = 3;

HSO_DEV bool arc9(unsigned m) {
  m |= m << 16;
  unsigned t = m & (m >> 1);
  t &= t >> 2;
  t &= t >> 4;   // runs of 8
  t &= m >> 8;   // runs of 9
  return t != 0;
}

__global__ void __launch_bounds__(FT_X * FT_Y) k_fast_score(const uint8_t* __restrict__ img, int w, int h, int threshold, int16_t* __restrict__ score) {
  __shared__ uint8_t tile[FT_Y + 2 * FT_H][FT_X + 2 * FT_H + 2];
  const int x0 = blockIdx.x * FT_X, y0 = blockIdx.y * FT_Y;
  for (int idx = threadIdx.y * FT_X + threadIdx.x; idx < (FT_Y + 2 * FT_H) * (FT_X + 2 * FT_H); idx += FT_X * FT_Y) {
    const int r = idx / (FT_X + 2 * FT_H), c = idx - r * (FT_X + 2 * FT_H);
    const int x = x0 + c - FT_H, y = y0 + r - FT_H;
    tile[r][c] = (x >= 0 && x < w && y >= 0 && y < h) ? img[(size_t)y * w + x] : 0;
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= w || y >= h) return;
  int out = -1;
  if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {  // faster_corner_9_sse.cpp:27-29 (y) and :31,:57,:233 (x)
    const int cx = threadIdx.x + FT_H, cy = threadIdx.y + FT_H;
    const int p = tile[cy][cx];
    int d[16];
    d[0] = tile[cy + 3][cx] - p;      d[1] = tile[cy + 3][cx + 1] - p;  d[2] = tile[cy + 2][cx + 2] - p;  d[3] = tile[cy + 1][cx + 3] - p;
    d[4] = tile[cy][cx + 3] - p;      d[5] = tile[cy - 1][cx + 3] - p;  d[6] = tile[cy - 2][cx + 2] - p;  d[7] = tile[cy - 3][cx + 1] - p;
    d[8] = tile[cy - 3][cx] - p;      d[9] = tile[cy - 3][cx - 1] - p;  d[10] = tile[cy - 2][cx - 2] - p; d[11] = tile[cy - 1][cx - 3] - p;
    d[12] = tile[cy][cx - 3] - p;     d[13] = tile[cy + 1][cx - 3] - p; d[14] = tile[cy + 2][cx - 2] - p; d[15] = tile[cy + 3][cx - 1] - p;
    auto corner_at = [&](int b) {
      unsigned bright = 0, dark = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        bright |= (d[i] > b ? 1u : 0u) << i;
        dark |= (d[i] < -b ? 1u : 0u) << i;
      }
      return arc9(bright) || arc9(dark);
    };
    if (corner_at(threshold)) {
      int lo = threshold, hi = 255;  // corner at lo, not a corner at hi (|d| <= 255)
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (corner_at(mid)) lo = mid; else hi = mid;
      }
      out = lo;
    }
  }
  score[(size_t)y * w + x] = (int16_t)out;
}

__global__ void __launch_bounds__(128) k_fast_rows(const int16_t* __restrict__ score, int w, int h, int border, uint32_t* __restrict__ rowbuf,
                                                     int* __restrict__ row_count) {
  __shared__ int warp_cnt[4];
  __shared__ int base_s;
  const int y = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int xb = 0; xb < w; xb += 128) {
    const int x = xb + threadIdx.x;
    bool keep = false;
    int s = -1;
    if (x < w) {
      s = score[(size_t)y * w + x];
      if (s >= 0) {
        keep = true;  // corners exist only in [3, w-3) x [3, h-3): all 8 neighbours are inside the map
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx)
            if ((dx | dy) != 0 && score[(size_t)(y + dy) * w + x + dx] >= s) keep = false;   // nonmax_3x3.cpp:58-104 (>=)
        if (x < border || x > w - border || y < border || y > h - border) keep = false;      // feature_detection.cpp:515
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int k = 0; k < warp; ++k) off += warp_cnt[k];
    if (keep) rowbuf[(size_t)y * w + off + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)x | ((uint32_t)s << 16);
    __syncthreads();
    if (threadIdx.x == 0) base_s += warp_cnt[0] + warp_cnt[1] + warp_cnt[2] + warp_cnt[3];
    __syncthreads();
  }
  if (threadIdx.x == 0) row_count[y] = base_s;
}

// hso::shiTomasiScore (src/vikit/vision.cpp:111-151): the three sums are integers below 2^24, hence exact in float in any order.
HSO_DEV float shi_tomasi(const uint8_t* img, int cols, int rows, int u, int v) {
  const int x_min = u - 4, x_max = u + 4, y_min = v - 4, y_max = v + 4;
  if (x_min < 1 || x_max >= cols - 1 || y_min < 1 || y_max >= rows - 1) return 0.f;
  int sxx = 0, syy = 0, sxy = 0;
  for (int y = y_min; y < y_max; ++y)
    for (int x = x_min; x < x_max; ++x) {
      const int dx = (int)img[(size_t)y * cols + x + 1] - (int)img[(size_t)y * cols + x - 1];
      const int dy = (int)img[(size_t)(y + 1) * cols + x] - (int)img[(size_t)(y - 1) * cols + x];
      sxx += dx * dx; syy += dy * dy; sxy += dx * dy;
    }
  const float dXX = (float)((double)(float)sxx / (2.0 * 64)), dYY = (float)((double)(float)syy / (2.0 * 64)), dXY = (float)((double)(float)sxy / (2.0 * 64));
  const float tr = dXX + dYY;
  return (float)(0.5 * (double)(tr - sqrtf(tr * tr - 4.f * (dXX * dYY - dXY * dXY))));
}

__global__ void __launch_bounds__(128) k_fast_gather(const uint8_t* __restrict__ img, int w, int h, const uint32_t* __restrict__ rowbuf,
                                                       const int* __restrict__ row_count, hso_corner* __restrict__ out, int cap, int* __restrict__ total) {
  __shared__ int red[4];
  const int y = blockIdx.x;
  int part = 0;
  for (int j = threadIdx.x; j < y; j += 128) part += row_count[j];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  const int off = red[0] + red[1] + red[2] + red[3];
  const int n = row_count[y];
  if (y == h - 1 && threadIdx.x == 0) *total = off + n;
  for (int k = threadIdx.x; k < n; k += 128) {
    if (off + k >= cap) break;
    const uint32_t v = rowbuf[(size_t)y * w + k];
    const int x = (int)(v & 0xffffu), s = (int)(v >> 16);
    hso_corner c;
    c.x = (int16_t)x; c.y = (int16_t)y; c.score = s; c.shi_tomasi = shi_tomasi(img, w, h, x, y);
    out[off + k] = c;
  }
}

cudaError_t launch_fast(const uint8_t* level_img, int w, int h, int threshold, int border, int16_t* score_map, uint32_t* rowbuf, int* row_count,
                        hso_corner* out_dev, int cap, int* total_dev, cudaStream_t stream, uint64_t* launches) {
  dim3 grid((w + FT_X - 1) / FT_X, (h + FT_Y - 1) / FT_Y), block(FT_X, FT_Y);
  k_fast_score<<<grid, block, 0, stream>>>(level_img, w, h, threshold, score_map);
  k_fast_rows<<<h, 128, 0, stream>>>(score_map, w, h, border, rowbuf, row_count);
  k_fast_gather<<<h, 128, 0, stream>>>(level_img, w, h, rowbuf, row_count, out_dev, cap, total_dev);
  *launches += 3;
  return cudaGetLastError();
}

}  // namespace hso
