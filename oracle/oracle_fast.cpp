// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of FeatureExtractor::fastDetectST (src/feature_detection.cpp:498-523):
// FAST-9 segment test (thirdparty/fast/src/faster_corner_9_sse.cpp:16-258, include/fast/corner_9.h), fast_corner_score_9
// (thirdparty/fast/src/fast_9_score.cpp:22-4681), fast_nonmax_3x3 (thirdparty/fast/src/nonmax_3x3.cpp:18-112), the 8-px border
// filter and hso::shiTomasiScore (src/vikit/vision.cpp:111-151). The generated decision trees are restated by their definition:
//   corner at barrier b  <=>  9 contiguous of the 16 Bresenham-circle pixels are all > p + b or all < p - b;
//   score = the largest barrier at which the pixel is still a corner (the generated code climbs b by the arc's min difference);
//   non-max survivor <=> no 8-neighbour corner has a score >= its own.
// Parity status: PINNED — this restatement is compared bit-for-bit with the real reference library compiled from
// /root/reference/thirdparty/fast (oracle/_ref/libfast_ref.so, tests/test_oracle_fast.py).
#include <cmath>
#include <cstdint>
#include <vector>

#include "hso_oracle.h"

namespace {

const int kCircle[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, -1},
                            {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};  // fast_9_score.cpp:4662-4679 (x, y*stride)

inline bool has_arc9(unsigned m) {  // 9 contiguous set bits in a circular 16-bit mask
  m |= m << 16;
  for (int s = 0; s < 16; ++s) if (((m >> s) & 0x1ffu) == 0x1ffu) return true;
  return false;
}

inline bool is_corner(const uint8_t* p, int stride, int b) {
  const int c = *p;
  unsigned bright = 0, dark = 0;
  for (int i = 0; i < 16; ++i) {
    const int v = p[kCircle[i][1] * stride + kCircle[i][0]];
    if (v > c + b) bright |= 1u << i;
    if (v < c - b) dark |= 1u << i;
  }
  return has_arc9(bright) || has_arc9(dark);
}

// src/vikit/vision.cpp:111-151
float shi_tomasi(const uint8_t* img, int cols, int rows, int stride, int u, int v) {
  float dXX = 0.0, dYY = 0.0, dXY = 0.0;
  const int halfbox_size = 4, box_size = 8, box_area = 64;
  const int x_min = u - halfbox_size, x_max = u + halfbox_size, y_min = v - halfbox_size, y_max = v + halfbox_size;
  if (x_min < 1 || x_max >= cols - 1 || y_min < 1 || y_max >= rows - 1) return 0.0;
  for (int y = y_min; y < y_max; ++y) {
    const uint8_t* ptr_left = img + stride * y + x_min - 1;
    const uint8_t* ptr_right = img + stride * y + x_min + 1;
    const uint8_t* ptr_top = img + stride * (y - 1) + x_min;
    const uint8_t* ptr_bottom = img + stride * (y + 1) + x_min;
    for (int x = 0; x < box_size; ++x, ++ptr_left, ++ptr_right, ++ptr_top, ++ptr_bottom) {
      float dx = *ptr_right - *ptr_left;
      float dy = *ptr_bottom - *ptr_top;
      dXX += dx * dx;
      dYY += dy * dy;
      dXY += dx * dy;
    }
  }
  dXX = dXX / (2.0 * box_area);
  dYY = dYY / (2.0 * box_area);
  dXY = dXY / (2.0 * box_area);
  return 0.5 * (dXX + dYY - std::sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY)));
}

}  // namespace

extern "C" {

// All FAST-9 corners in raster order with their scores (the reference's fastCorners9 / scores9), x in [3, w-3), y in [3, h-3).
int orc_fast9_corners(const uint8_t* img, int w, int h, int stride, int threshold, int16_t* xy, int32_t* scores, int cap) {
  int n = 0;
  if (w < 7 || h < 7) return 0;
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      const uint8_t* p = img + (size_t)y * stride + x;
      if (!is_corner(p, stride, threshold)) continue;
      int b = threshold;
      while (b < 255 && is_corner(p, stride, b + 1)) ++b;
      if (n < cap) { xy[2 * n] = (int16_t)x; xy[2 * n + 1] = (int16_t)y; scores[n] = b; }
      ++n;
    }
  return n;
}

// FeatureExtractor::fastDetectST after the detector: non-max suppression, border filter, Shi-Tomasi score. Output in raster order.
int orc_fast_detect(const uint8_t* img, int w, int h, int stride, int threshold, int border, orc_corner* out, int cap) {
  std::vector<int16_t> xy((size_t)2 * w * h);
  std::vector<int32_t> sc((size_t)w * h);
  const int n = orc_fast9_corners(img, w, h, stride, threshold, xy.data(), sc.data(), w * h);
  std::vector<int32_t> map((size_t)w * h, -1);
  for (int i = 0; i < n; ++i) map[(size_t)xy[2 * i + 1] * w + xy[2 * i]] = sc[i];
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const int x = xy[2 * i], y = xy[2 * i + 1], s = sc[i];
    bool keep = true;
    for (int dy = -1; dy <= 1 && keep; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        if (!dx && !dy) continue;
        if (map[(size_t)(y + dy) * w + x + dx] >= s) { keep = false; break; }  // corners live in [3, w-3) x [3, h-3): neighbours exist
      }
    if (!keep) continue;
    if (x < border || x > w - border || y < border || y > h - border) continue;  // feature_detection.cpp:515
    if (m < cap) { out[m].x = (int16_t)x; out[m].y = (int16_t)y; out[m].score = s; out[m].shi_tomasi = shi_tomasi(img, w, h, stride, x, y); }
    ++m;
  }
  return m;
}

}  // extern "C"
