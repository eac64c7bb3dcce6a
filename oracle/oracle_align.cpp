// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the live (float) overloads of
// feature_alignment::align1D / align2D (src/feature_alignment.cpp:164-308, 464-605) and of the direct part of
// hso::Matcher (src/matcher.cpp:74-85 getBestSearchLevel, :120-155 warpAffine(float), :226-238
// createPatchFromPatchWithBorder, :310-375 findMatchDirect tail, :379-404 checkNCC, :406-440 checkNormal;
// include/hso/vikit/vision.h:49-65 interpolateMat_8u). Float throughout, like the reference.
// Parity status: unpinned by the reference (no tests/golden vectors exist for these functions).
#include <cmath>
#include <cstring>

#include "hso_oracle.h"
#include "oracle_math.hpp"

using namespace orc;

namespace {

// `int u_r = floor(u)` on a NaN is undefined behaviour; x86-64 (cvttsd2si) yields INT_MIN, which fails the
// `u_r < halfpatch_size_` test and leaves the loop before the explicit isnan() check is ever reached
// (feature_alignment.cpp:531-538). The restatement fixes that platform behaviour.
inline int floor_to_int_x86(float v) { return std::isnan(v) ? INT32_MIN : (int)std::floor(v); }

// include/hso/vikit/vision.h:49-65
inline float interpolateMat_8u(const uint8_t* data, int stride, float u, float v) {
  int x = floor(u);
  int y = floor(v);
  float subpix_x = u - x;
  float subpix_y = v - y;
  float w00 = (1.0f - subpix_x) * (1.0f - subpix_y);
  float w01 = (1.0f - subpix_x) * subpix_y;
  float w10 = subpix_x * (1.0f - subpix_y);
  float w11 = 1.0f - w00 - w01 - w10;
  const uint8_t* ptr = data + y * stride + x;
  return w00 * ptr[0] + w01 * ptr[stride] + w10 * ptr[1] + w11 * ptr[stride + 1];
}

}  // namespace

extern "C" {

// src/feature_alignment.cpp:464-605
int orc_align2d(const uint8_t* cur_img, int cols, int rows, int stride, const float* ref_patch_with_border, const float* ref_patch,
                int n_iter, double px_io[2], float* cur_patch) {
  const int halfpatch_size_ = 4, patch_size_ = 8, patch_area_ = 64;
  bool converged = false;
  float ref_patch_dx[patch_area_], ref_patch_dy[patch_area_], grad_weight[patch_area_];
  float H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int ref_step = patch_size_ + 2;
  {
    float *it_dx = ref_patch_dx, *it_dy = ref_patch_dy, *it_weight = grad_weight;
    float J[3];
    for (int y = 0; y < patch_size_; ++y) {
      const float* it = ref_patch_with_border + (y + 1) * ref_step + 1;
      for (int x = 0; x < patch_size_; ++x, ++it, ++it_dx, ++it_dy, ++it_weight) {
        J[0] = 0.5 * (it[1] - it[-1]);
        J[1] = 0.5 * (it[ref_step] - it[-ref_step]);
        J[2] = 1.;
        *it_dx = J[0];
        *it_dy = J[1];
        *it_weight = sqrtf(250.0 / (250.0 + (J[0] * J[0] + J[1] * J[1])));
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) H[r * 3 + c] += (J[r] * J[c]) * (*it_weight);
      }
    }
  }
  for (int i = 0; i < 3; i++) H[i * 3 + i] *= (1 + 0.001);
  float Hinv[9];
  inv3f(H, Hinv);

  float u = px_io[0];
  float v = px_io[1];
  const float min_update_squared = 0.03 * 0.03;
  const int cur_step = stride;
  float mean_diff = 0;
  float chi2 = 0;
  float update[3] = {0, 0, 0};
  float Jres[3] = {0, 0, 0};
  for (int iter = 0; iter < n_iter; ++iter) {
    float* cur_patch_ptr = cur_patch;
    int u_r = floor_to_int_x86(u);
    int v_r = floor_to_int_x86(v);
    if (u_r < halfpatch_size_ || v_r < halfpatch_size_ || u_r >= cols - halfpatch_size_ || v_r >= rows - halfpatch_size_) break;
    if (std::isnan(u) || std::isnan(v)) return 0;
    float subpix_x = u - u_r;
    float subpix_y = v - v_r;
    float wTL = (1.0 - subpix_x) * (1.0 - subpix_y);
    float wTR = subpix_x * (1.0 - subpix_y);
    float wBL = (1.0 - subpix_x) * subpix_y;
    float wBR = subpix_x * subpix_y;
    const float *it_ref = ref_patch, *it_ref_dx = ref_patch_dx, *it_ref_dy = ref_patch_dy, *it_weight = grad_weight;
    float new_chi2 = 0.0;
    Jres[0] = Jres[1] = Jres[2] = 0;
    for (int y = 0; y < patch_size_; ++y) {
      const uint8_t* it = cur_img + (v_r + y - halfpatch_size_) * cur_step + u_r - halfpatch_size_;
      for (int x = 0; x < patch_size_; ++x, ++it, ++it_ref, ++it_ref_dx, ++it_ref_dy, ++it_weight) {
        float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[cur_step] + wBR * it[cur_step + 1];
        float res = search_pixel - (*it_ref) + mean_diff;
        Jres[0] -= res * (*it_ref_dx) * (*it_weight);
        Jres[1] -= res * (*it_ref_dy) * (*it_weight);
        Jres[2] -= res * (*it_weight);
        new_chi2 += res * res * (*it_weight);
        if (cur_patch != NULL) { *cur_patch_ptr = search_pixel; ++cur_patch_ptr; }
      }
    }
    chi2 = new_chi2;
    for (int r = 0; r < 3; ++r) update[r] = Hinv[r * 3 + 0] * Jres[0] + Hinv[r * 3 + 1] * Jres[1] + Hinv[r * 3 + 2] * Jres[2];
    u += update[0];
    v += update[1];
    mean_diff += update[2];
    if (update[0] * update[0] + update[1] * update[1] < min_update_squared) { converged = true; break; }
  }
  if (chi2 > 1000 * patch_area_) converged = false;
  px_io[0] = u;
  px_io[1] = v;
  return converged ? 1 : 0;
}

// src/feature_alignment.cpp:164-308
int orc_align1d(const uint8_t* cur_img, int cols, int rows, int stride, const float dir[2], const float* ref_patch_with_border,
                const float* ref_patch, int n_iter, double px_io[2], double* h_inv_out, float* cur_patch) {
  const int halfpatch_size_ = 4, patch_size = 8, patch_area = 64;
  bool converged = false;
  float ref_patch_dv[patch_area], grad_weight[patch_area];
  float H[4] = {0, 0, 0, 0};
  const int ref_step = patch_size + 2;
  {
    float *it_dv = ref_patch_dv, *it_weight = grad_weight;
    float J[2];
    for (int y = 0; y < patch_size; ++y) {
      const float* it = ref_patch_with_border + (y + 1) * ref_step + 1;
      for (int x = 0; x < patch_size; ++x, ++it, ++it_dv, ++it_weight) {
        J[0] = 0.5 * (dir[0] * (it[1] - it[-1]) + dir[1] * (it[ref_step] - it[-ref_step]));
        J[1] = 1.;
        *it_dv = J[0];
        *it_weight = sqrtf(250.0 / (250.0 + J[0] * J[0]));
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 2; ++c) H[r * 2 + c] += (J[r] * J[c]) * (*it_weight);
      }
    }
  }
  for (int i = 0; i < 2; i++) H[i * 2 + i] *= (1 + 0.001);
  double h_inv = 1.0 / H[0] * patch_size * patch_size;
  if (h_inv_out) *h_inv_out = h_inv;
  float Hinv[4];
  inv2f(H, Hinv);
  float mean_diff = 0;
  float u = px_io[0];
  float v = px_io[1];
  const float min_update_squared = 0.01 * 0.01;
  const int cur_step = stride;
  float chi2 = 0;
  float update[2] = {0, 0};
  float Jres[2] = {0, 0};
  for (int iter = 0; iter < n_iter; ++iter) {
    float* cur_patch_ptr = cur_patch;
    int u_r = floor_to_int_x86(u);
    int v_r = floor_to_int_x86(v);
    if (u_r < halfpatch_size_ || v_r < halfpatch_size_ || u_r >= cols - halfpatch_size_ || v_r >= rows - halfpatch_size_) break;
    if (std::isnan(u) || std::isnan(v)) return 0;
    float subpix_x = u - u_r;
    float subpix_y = v - v_r;
    float wTL = (1.0 - subpix_x) * (1.0 - subpix_y);
    float wTR = subpix_x * (1.0 - subpix_y);
    float wBL = (1.0 - subpix_x) * subpix_y;
    float wBR = subpix_x * subpix_y;
    const float *it_ref = ref_patch, *it_ref_dv = ref_patch_dv, *it_weight = grad_weight;
    float new_chi2 = 0.0;
    Jres[0] = Jres[1] = 0;
    for (int y = 0; y < patch_size; ++y) {
      const uint8_t* it = cur_img + (v_r + y - halfpatch_size_) * cur_step + u_r - halfpatch_size_;
      for (int x = 0; x < patch_size; ++x, ++it, ++it_ref, ++it_ref_dv, ++it_weight) {
        float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[cur_step] + wBR * it[cur_step + 1];
        float res = search_pixel - *it_ref + mean_diff;
        Jres[0] -= res * (*it_ref_dv) * (*it_weight);
        Jres[1] -= res * (*it_weight);
        new_chi2 += res * res * (*it_weight);
        if (cur_patch != NULL) { *cur_patch_ptr = search_pixel; ++cur_patch_ptr; }
      }
    }
    chi2 = new_chi2;
    update[0] = Hinv[0] * Jres[0] + Hinv[1] * Jres[1];
    update[1] = Hinv[2] * Jres[0] + Hinv[3] * Jres[1];
    u += update[0] * dir[0];
    v += update[0] * dir[1];
    mean_diff += update[1];
    if (update[0] * update[0] < min_update_squared) { converged = true; break; }
  }
  if (chi2 > 1000 * patch_area) converged = false;
  px_io[0] = u;
  px_io[1] = v;
  return converged ? 1 : 0;
}

// src/matcher.cpp:74-85
int orc_get_best_search_level(const double A[4], int max_level) {
  int search_level = 0;
  double D = A[0] * A[3] - A[1] * A[2];
  while (D > 3.0 && search_level < max_level) { search_level += 1; D *= 0.25; }
  return search_level;
}

// src/matcher.cpp:120-155 (float overload). On a NaN warp the reference returns without writing the patch (stack
// garbage downstream); the restatement zero-fills, which makes the downstream NCC test fail deterministically.
void orc_warp_affine(const double A[4], const uint8_t* img_ref, int cols, int rows, const double px_ref[2], int level_ref,
                     int search_level, int halfpatch_size, float* patch) {
  const int patch_size = halfpatch_size * 2;
  const double det = A[0] * A[3] - A[1] * A[2];
  const double invdet = 1.0 / det;  // Eigen 2x2 inverse: adjugate * (1/det)
  const float Ai[4] = {(float)(A[3] * invdet), (float)(-A[1] * invdet), (float)(-A[2] * invdet), (float)(A[0] * invdet)};
  if (std::isnan(Ai[0])) { std::memset(patch, 0, sizeof(float) * patch_size * patch_size); return; }
  float* patch_ptr = patch;
  const float px_ref_pyr[2] = {(float)(px_ref[0] / (1 << level_ref)), (float)(px_ref[1] / (1 << level_ref))};
  const float scaleTarget = (1 << search_level);
  for (int y = 0; y < patch_size; ++y)
    for (int x = 0; x < patch_size; ++x, ++patch_ptr) {
      float pp[2] = {(float)(x - halfpatch_size), (float)(y - halfpatch_size)};
      pp[0] *= scaleTarget;
      pp[1] *= scaleTarget;
      const float px0 = (Ai[0] * pp[0] + Ai[1] * pp[1]) + px_ref_pyr[0];
      const float px1 = (Ai[2] * pp[0] + Ai[3] * pp[1]) + px_ref_pyr[1];
      if (px0 < 0 || px1 < 0 || px0 >= cols - 1 || px1 >= rows - 1) *patch_ptr = 0;
      else *patch_ptr = interpolateMat_8u(img_ref, cols, px0, px1);
    }
}

// src/matcher.cpp:379-404
int orc_check_ncc(const float* patch1, const float* patch2, float thresh) {
  const int NCC_area = 64;
  float mean1 = 0, mean2 = 0;
  for (int i = 0; i < NCC_area; ++i) { mean1 += patch1[i]; mean2 += patch2[i]; }
  mean1 /= NCC_area;
  mean2 /= NCC_area;
  float numerator = 0, demoniator1 = 0, demoniator2 = 0;
  for (int i = 0; i < NCC_area; i++) {
    float patch1_mean = patch1[i] - mean1;
    float patch2_mean = patch2[i] - mean2;
    numerator += patch1_mean * patch2_mean;
    demoniator1 += patch1_mean * patch1_mean;
    demoniator2 += patch2_mean * patch2_mean;
  }
  return (numerator / (std::sqrt(demoniator1 * demoniator2) + 1e-12)) > thresh;
}

// src/matcher.cpp:406-440
int orc_check_normal(const int16_t* sobx, const int16_t* soby, int cols, const double pxLevel[2], const double normal[2], float thresh) {
  float uf = pxLevel[0];
  float vf = pxLevel[1];
  int ui = floorf(pxLevel[0]);
  int vi = floorf(pxLevel[1]);
  float subpix_x = uf - ui;
  float subpix_y = vf - vi;
  float wTL = (1.0 - subpix_x) * (1.0 - subpix_y);
  float wTR = subpix_x * (1.0 - subpix_y);
  float wBL = (1.0 - subpix_x) * subpix_y;
  float wBR = 1.0 - wTL - wTR - wBL;
  const size_t o = (size_t)vi * cols + ui;
  short gx00 = sobx[o], gx10 = sobx[o + 1], gx01 = sobx[o + cols], gx11 = sobx[o + cols + 1];
  short gy00 = soby[o], gy10 = soby[o + 1], gy01 = soby[o + cols], gy11 = soby[o + cols + 1];
  double nx = wTL * (double)gx00 + wTR * (double)gx10 + wBL * (double)gx01 + wBR * (double)gx11;
  double ny = wTL * (double)gy00 + wTR * (double)gy10 + wBL * (double)gy01 + wBR * (double)gy11;
  double nn = std::sqrt(nx * nx + ny * ny);
  nx /= nn;
  ny /= nn;
  return (normal[0] * nx + normal[1] * ny) > thresh;
}

// Tail of Matcher::findMatchDirect — src/matcher.cpp:310-375 (everything after getWarpMatrixAffine/getBestSearchLevel,
// whose inputs need cam2world and the map and therefore stay with the host caller).
void orc_match_direct_batch(int M, const orc_align_job* jobs, const uint8_t* const* ref_levels, const uint8_t* const* cur_levels,
                            const int* lw, const int* lh, const int16_t* const* cur_sobx, const int16_t* const* cur_soby,
                            int align_max_iter, orc_align_result* out) {
  const int patch_size_ = 8, halfpatch_size_ = 4;
  for (int m = 0; m < M; ++m) {
    const orc_align_job& jb = jobs[m];
    orc_align_result& rs = out[m];
    float patch_with_border_f_temp[100], patch_with_border_f_[100], patch_f_[64];
    const int rl = jb.ref_level, sl = jb.search_level;
    orc_warp_affine(jb.A_cur_ref, ref_levels[rl], lw[rl], lh[rl], jb.px_ref, rl, sl, halfpatch_size_ + 1, patch_with_border_f_temp);
    if (jb.scale_patch) {
      for (int i = 0; i < 100; ++i) patch_with_border_f_[i] = patch_with_border_f_temp[i] * jb.exposure_rat;
    } else {
      std::memcpy(patch_with_border_f_, patch_with_border_f_temp, sizeof patch_with_border_f_);
    }
    for (int y = 1; y < patch_size_ + 1; ++y)
      for (int x = 0; x < patch_size_; ++x) patch_f_[(y - 1) * patch_size_ + x] = patch_with_border_f_[y * (patch_size_ + 2) + 1 + x];
    double px_scaled[2] = {jb.px_cur[0] / (1 << sl), jb.px_cur[1] / (1 << sl)};
    const double px_scaled_orig[2] = {px_scaled[0], px_scaled[1]};
    float patchNCC[64];
    std::memset(patchNCC, 0, sizeof patchNCC);
    int alignResult;
    rs.h_inv = 0;
    if (jb.type == 1) {
      double dir_cur[2] = {jb.A_cur_ref[0] * jb.grad[0] + jb.A_cur_ref[1] * jb.grad[1], jb.A_cur_ref[2] * jb.grad[0] + jb.A_cur_ref[3] * jb.grad[1]};
      double n = std::sqrt(dir_cur[0] * dir_cur[0] + dir_cur[1] * dir_cur[1]);
      dir_cur[0] /= n;
      dir_cur[1] /= n;
      const float dirf[2] = {(float)dir_cur[0], (float)dir_cur[1]};
      alignResult = orc_align1d(cur_levels[sl], lw[sl], lh[sl], lw[sl], dirf, patch_with_border_f_, patch_f_, align_max_iter, px_scaled,
                                &rs.h_inv, patchNCC);
      rs.align_converged = alignResult;
      if (alignResult) alignResult = orc_check_normal(cur_sobx[sl], cur_soby[sl], lw[sl], px_scaled, dir_cur, 0.86 /* Config::edgeLetCosAngle, config.cpp:58 */);
    } else {
      alignResult = orc_align2d(cur_levels[sl], lw[sl], lh[sl], lw[sl], patch_with_border_f_, patch_f_, align_max_iter, px_scaled, patchNCC);
      rs.align_converged = alignResult;
    }
    if (alignResult) alignResult = orc_check_ncc(patch_f_, patchNCC, jb.ncc_thresh > 0.f ? jb.ncc_thresh : 0.7);
    if (alignResult) {
      double dx = px_scaled_orig[0] - px_scaled[0], dy = px_scaled_orig[1] - px_scaled[1];
      alignResult = std::sqrt(dx * dx + dy * dy) < 20;
    }
    rs.px_cur[0] = px_scaled[0] * (1 << sl);
    rs.px_cur[1] = px_scaled[1] * (1 << sl);
    rs.ok = alignResult;
  }
}

}  // extern "C"
