// ORACLE — TEST INFRASTRUCTURE ONLY. Thin extern "C" wrapper around the REAL reference code of row N2:
// the vendored FAST library of the reference, compiled from the sources where they lie (/root/reference/thirdparty/fast/src:
// faster_corner_9_sse.cpp, fast_9.cpp, fast_9_score.cpp, nonmax_3x3.cpp) into oracle/_ref/libfast_ref.so by oracle/Makefile.
// This file contains no reference code; it only calls fast::fast_corner_detect_9_sse2 / fast_corner_score_9 / fast_nonmax_3x3 in the
// order FeatureExtractor::fastDetectST does (src/feature_detection.cpp:498-523).
#include <fast/fast.h>

#include <cstdint>
#include <vector>

extern "C" {

// Returns the number of detected corners (before non-max suppression); fills up to cap entries of xy (x,y interleaved) and scores,
// and up to cap indices of the non-max survivors (n_nonmax receives their count).
int ref_fast9_detect(const uint8_t* img, int w, int h, int stride, int threshold, int16_t* xy, int32_t* scores, int32_t* nonmax_idx, int cap,
                     int* n_nonmax) {
  std::vector<fast::fast_xy> corners;
  fast::fast_corner_detect_9_sse2((fast::fast_byte*)img, w, h, stride, (short)threshold, corners);
  std::vector<int> sc, nm;
  fast::fast_corner_score_9((fast::fast_byte*)img, stride, corners, threshold, sc);
  fast::fast_nonmax_3x3(corners, sc, nm);
  const int n = (int)corners.size();
  for (int i = 0; i < n && i < cap; ++i) { xy[2 * i] = corners[i].x; xy[2 * i + 1] = corners[i].y; scores[i] = sc[i]; }
  for (int i = 0; i < (int)nm.size() && i < cap; ++i) nonmax_idx[i] = nm[i];
  if (n_nonmax) *n_nonmax = (int)nm.size();
  return n;
}

}  // extern "C"
