// ORACLE — test infrastructure, never linked into or called by the product (hso_b200/).
// CPU restatement of row N3: DepthFilter::observeDepthRow (src/depth_filter.cpp:580-675) for a list of seeds against one active frame:
// visibility test, inverse-depth search interval, Matcher::doLineStereo (src/matcher.cpp:802-1049: warp, exposure scaling, epipolar segment
// in the search level, ZMNCC_F scan (include/hso/vikit/patch_score.h:268-305) over warp::createPatch samples (src/matcher.cpp:159-196),
// ambiguity test, Matcher::KLTLimited1D / KLTLimited2D refinement (:1296-1606), checkNormal, checkNCC, depthFromTriangulation (:242-255)),
// DepthFilter::computeTau (:539-555) and DepthFilter::updateSeed (:528-537). Line-faithful loops, same float/double mix, same quirks.
//
// One defined-where-the-reference-is-undefined case: KLTLimited2D / KLTLimited1D leave `targetPatch` (patch2D, an uninitialised stack
// array in doLineStereo, :987) unwritten when their first iteration breaks at the image border; the oracle and the CUDA path start it at 0.
#include <cmath>
#include <cstring>
#include <vector>

#include "hso_oracle.h"
#include "oracle_math.hpp"

using namespace orc;

namespace {

struct Img { const uint8_t* data; int cols, rows; };

// warp::createPatch float overload — src/matcher.cpp:159-196
void create_patch(float* patch, const double px_scaled[2], const Img& im, int halfpatch_size) {
  const int patch_size = halfpatch_size * 2;
  const int stride = im.cols;
  const float u_cur = px_scaled[0], v_cur = px_scaled[1];
  const int ui = floorf(u_cur), vi = floorf(v_cur);
  const float subpix_u_ref = u_cur - ui, subpix_v_ref = v_cur - vi;
  const float w_ref_tl = (1.0 - subpix_u_ref) * (1.0 - subpix_v_ref);
  const float w_ref_tr = subpix_u_ref * (1.0 - subpix_v_ref);
  const float w_ref_bl = (1.0 - subpix_u_ref) * subpix_v_ref;
  const float w_ref_br = 1.0 - w_ref_tl - w_ref_tr - w_ref_bl;
  float* patch_ptr = patch;
  for (int y = 0; y < patch_size; ++y) {
    const uint8_t* p = im.data + (vi - halfpatch_size + y) * stride + (ui - halfpatch_size);
    for (int x = 0; x < patch_size; ++x, ++patch_ptr, ++p)
      *patch_ptr = w_ref_tl * p[0] + w_ref_tr * p[1] + w_ref_bl * p[stride] + w_ref_br * p[stride + 1];
  }
}

// patch_score::ZMNCC_F<4> — include/hso/vikit/patch_score.h:268-305
struct ZMNCC {
  const float* host;
  float hostMean = 0;
  explicit ZMNCC(const float* ref) : host(ref) {
    for (int r = 0; r < 64; r++) hostMean += host[r];
    hostMean /= 64;
  }
  float score(const float* target) const {
    float targetMean = 0;
    for (int r = 0; r < 64; r++) targetMean += target[r];
    targetMean /= 64;
    float numerator = 0, demoniator1 = 0, demoniator2 = 0;
    for (int i = 0; i < 64; i++) {
      const float h = host[i] - hostMean;
      const float t = target[i] - targetMean;
      numerator += h * t;
      demoniator1 += h * h;
      demoniator2 += t * t;
    }
    return (numerator / (std::sqrt(demoniator1 * demoniator2) + 1e-12));
  }
};

// Matcher::KLTLimited2D — src/matcher.cpp:1296-1450
bool klt_limited_2d(const Img& target, const float* hostPatchWithBorder, const float* hostPatch, int n_iter, double px[2], float* targetPatch) {
  const int halfPatchSize = 4, patchSize = 8, patchArea = 64;
  float host_dx[64], host_dy[64], grad_weight[64];
  float H[9] = {0};
  const int hostStep = patchSize + 2;
  int k = 0;
  for (int y = 0; y < patchSize; ++y) {
    const float* it = hostPatchWithBorder + (y + 1) * hostStep + 1;
    for (int x = 0; x < patchSize; ++x, ++it, ++k) {
      float J[3];
      J[0] = 0.5 * (it[1] - it[-1]);
      J[1] = 0.5 * (it[hostStep] - it[-hostStep]);
      J[2] = 1;
      host_dx[k] = J[0];
      host_dy[k] = J[1];
      grad_weight[k] = sqrtf(250.0 / (250.0 + (J[0] * J[0] + J[1] * J[1])));
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) H[3 * a + b] += J[a] * J[b] * grad_weight[k];
    }
  }
  for (int i = 0; i < 3; i++) H[4 * i] *= (1 + 0.001);
  float Hinv[9];
  inv3f(H, Hinv);
  float mean_diff = 0;
  float bestU = px[0], bestV = px[1];
  const int cur_step = target.cols;
  float bestEnergy = 1e8;
  float step[3] = {0, 0, 0}, stepBack[3] = {0, 0, 0}, Jres[3] = {0, 0, 0};
  float uBak = bestU, vBak = bestV, meanBak = mean_diff;
  for (int iter = 0; iter < n_iter; ++iter) {
    float* cur_patch_ptr = targetPatch;
    const int u_r = floor(bestU), v_r = floor(bestV);
    if (u_r < halfPatchSize || v_r < halfPatchSize || u_r >= target.cols - halfPatchSize || v_r >= target.rows - halfPatchSize) break;
    if (std::isnan(bestU) || std::isnan(bestV)) return false;
    const float subpix_x = bestU - u_r, subpix_y = bestV - v_r;
    const float wTL = (1.0 - subpix_x) * (1.0 - subpix_y);
    const float wTR = subpix_x * (1.0 - subpix_y);
    const float wBL = (1.0 - subpix_x) * subpix_y;
    const float wBR = subpix_x * subpix_y;
    float energy = 0.0;
    Jres[0] = Jres[1] = Jres[2] = 0;
    int q = 0;
    for (int y = 0; y < patchSize; ++y) {
      const uint8_t* it = target.data + (v_r + y - halfPatchSize) * cur_step + u_r - halfPatchSize;
      for (int x = 0; x < patchSize; ++x, ++it, ++q, ++cur_patch_ptr) {
        const float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[cur_step] + wBR * it[cur_step + 1];
        if (!std::isfinite(search_pixel)) { energy += 1e5; continue; }
        const float res = search_pixel - hostPatch[q] + mean_diff;
        Jres[0] -= res * host_dx[q] * grad_weight[q];
        Jres[1] -= res * host_dy[q] * grad_weight[q];
        Jres[2] -= res * grad_weight[q];
        energy += res * res * grad_weight[q];
        *cur_patch_ptr = search_pixel;
      }
    }
    if (energy > bestEnergy) {
      for (int a = 0; a < 3; ++a) stepBack[a] *= 0.5;
      bestU = uBak + stepBack[0];
      bestV = vBak + stepBack[1];
      mean_diff = meanBak + stepBack[2];
    } else {
      for (int a = 0; a < 3; ++a) step[a] = Hinv[3 * a] * Jres[0] + Hinv[3 * a + 1] * Jres[1] + Hinv[3 * a + 2] * Jres[2];
      if (step[0] < -0.5) step[0] = -0.5; else if (step[0] > 0.5) step[0] = 0.5;
      if (step[1] < -0.5) step[1] = -0.5; else if (step[1] > 0.5) step[1] = 0.5;
      if (!std::isfinite(step[0])) step[0] = step[1] = step[2] = 0;
      uBak = bestU; vBak = bestV; meanBak = mean_diff;
      for (int a = 0; a < 3; ++a) stepBack[a] = step[a];
      bestU += step[0];
      bestV += step[1];
      mean_diff += step[2];
      bestEnergy = energy;
    }
    if (stepBack[0] * stepBack[1] < 0.01 * 0.01) break;  // quirk: a product, negative when the components differ in sign
  }
  px[0] = bestU; px[1] = bestV;
  if (bestEnergy > 650 * patchArea) return false;
  return true;
}

// Matcher::KLTLimited1D — src/matcher.cpp:1454-1606
bool klt_limited_1d(const Img& target, const float* hostPatchWithBorder, const float* hostPatch, int n_iter, double px[2], const double direct[2],
                    float* targetPatch) {
  const int halfPatchSize = 4, patchSize = 8, patchArea = 64;
  float host_d[64], grad_weight[64];
  float H[4] = {0};
  const int hostStep = patchSize + 2;
  int k = 0;
  for (int y = 0; y < patchSize; ++y) {
    const float* it = hostPatchWithBorder + (y + 1) * hostStep + 1;
    for (int x = 0; x < patchSize; ++x, ++it, ++k) {
      float J[2];
      J[0] = 0.5 * (direct[0] * (it[1] - it[-1]) + direct[1] * (it[hostStep] - it[-hostStep]));
      J[1] = 1;
      host_d[k] = J[0];
      grad_weight[k] = sqrtf(250.0 / (250.0 + (J[0] * J[0])));
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) H[2 * a + b] += J[a] * J[b] * grad_weight[k];
    }
  }
  for (int i = 0; i < 2; i++) H[3 * i] *= (1 + 0.001);
  float Hinv[4];
  inv2f(H, Hinv);
  float mean_diff = 0;
  float bestU = px[0], bestV = px[1];
  const int cur_step = target.cols;
  float bestEnergy = 1e8;
  float step[2] = {0, 0}, stepBack[2] = {0, 0}, Jres[2] = {0, 0};
  float uBak = bestU, vBak = bestV, meanBak = mean_diff;
  for (int iter = 0; iter < n_iter; ++iter) {
    float* cur_patch_ptr = targetPatch;
    const int u_r = floor(bestU), v_r = floor(bestV);
    if (u_r < halfPatchSize || v_r < halfPatchSize || u_r >= target.cols - halfPatchSize || v_r >= target.rows - halfPatchSize) break;
    if (std::isnan(bestU) || std::isnan(bestV)) return false;
    const float subpix_x = bestU - u_r, subpix_y = bestV - v_r;
    const float wTL = (1.0 - subpix_x) * (1.0 - subpix_y);
    const float wTR = subpix_x * (1.0 - subpix_y);
    const float wBL = (1.0 - subpix_x) * subpix_y;
    const float wBR = subpix_x * subpix_y;
    float energy = 0.0;
    Jres[0] = Jres[1] = 0;
    int q = 0;
    for (int y = 0; y < patchSize; ++y) {
      const uint8_t* it = target.data + (v_r + y - halfPatchSize) * cur_step + u_r - halfPatchSize;
      for (int x = 0; x < patchSize; ++x, ++it, ++q) {
        const float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[cur_step] + wBR * it[cur_step + 1];
        if (!std::isfinite(search_pixel)) { energy += 1e5; continue; }
        const float res = search_pixel - hostPatch[q] + mean_diff;
        Jres[0] -= res * host_d[q] * grad_weight[q];
        Jres[1] -= res * grad_weight[q];
        energy += res * res * grad_weight[q];
        if (targetPatch != NULL) { *cur_patch_ptr = search_pixel; ++cur_patch_ptr; }
      }
    }
    if (energy > bestEnergy) {
      stepBack[0] *= 0.5; stepBack[1] *= 0.5;
      bestU = uBak + stepBack[0] * direct[0];
      bestV = vBak + stepBack[0] * direct[1];
      mean_diff = meanBak + stepBack[1];
    } else {
      step[0] = Hinv[0] * Jres[0] + Hinv[1] * Jres[1];
      step[1] = Hinv[2] * Jres[0] + Hinv[3] * Jres[1];
      if (step[0] < -0.5) step[0] = -0.5; else if (step[0] > 0.5) step[0] = 0.5;
      if (!std::isfinite(step[0])) step[0] = step[1] = 0;
      uBak = bestU; vBak = bestV; meanBak = mean_diff;
      stepBack[0] = step[0]; stepBack[1] = step[1];
      bestU += step[0] * direct[0];
      bestV += step[0] * direct[1];
      mean_diff += step[1];
      bestEnergy = energy;
    }
    if (fabsf(stepBack[0]) < 0.01) break;
  }
  px[0] = bestU; px[1] = bestV;
  if (bestEnergy > 650 * patchArea) return false;
  return true;
}

// depthFromTriangulation — src/matcher.cpp:242-255
bool depth_from_triangulation(const SE3& T_search_ref, const V3& f_ref, const V3& f_cur, double& depth) {
  const V3 a0 = T_search_ref.rotation() * f_ref;
  const V3 a1 = f_cur;
  const double AtA[4] = {a0.dot(a0), a0.dot(a1), a1.dot(a0), a1.dot(a1)};
  const double det = AtA[0] * AtA[3] - AtA[1] * AtA[2];
  if (det < 0.000001) return false;
  // Matrix2d::inverse(): adjugate / determinant
  const double inv[4] = {AtA[3] / det, -AtA[1] / det, -AtA[2] / det, AtA[0] / det};
  const double Att[2] = {a0.dot(T_search_ref.t), a1.dot(T_search_ref.t)};
  const double d0 = -(inv[0] * Att[0] + inv[1] * Att[1]);
  depth = std::fabs(d0);
  return true;
}

inline bool in_frame(const orc_cam* cam, int ox, int oy, int boundary, int level) {  // camera.h:85-89
  return ox >= boundary && ox < cam->width / (1 << level) - boundary && oy >= boundary && oy < cam->height / (1 << level) - boundary;
}

struct LineStereoOut { int search_level; int epl_start[2], epl_end[2]; double px_cur[2]; };

// Matcher::doLineStereo — src/matcher.cpp:802-1049
int do_line_stereo(const orc_cam* cam, const SE3& T_cur_ref, const orc_seed_obs& s, const uint8_t* const* ref_levels, const uint8_t* const* cur_levels,
                   const int* lw, const int* lh, const int16_t* const* cur_sobx, const int16_t* const* cur_soby, int max_search_level, int align_max_iter,
                   double min_idepth, double prior_idepth, double max_idepth, double& result_depth, LineStereoOut& o) {
  const int halfpatch_size_ = 4, patch_size_ = 8;
  double rt[12];
  T_cur_ref.to_rt(rt);
  double A[4];
  orc_get_warp_matrix_affine(cam, s.px, s.f, prior_idepth, rt, s.level, A);
  const int search_level = orc_get_best_search_level(A, max_search_level);
  o.search_level = search_level;
  float patch_with_border_f[100], patch_f[64];
  orc_warp_affine(A, ref_levels[s.level], lw[s.level], lh[s.level], s.px, s.level, search_level, halfpatch_size_ + 1, patch_with_border_f);
  const float exposure_rat = s.exposure_rat;
  if (fabsf(exposure_rat * 128 - 128) > 30.0f)
    for (int i = 0; i < 100; ++i) patch_with_border_f[i] = patch_with_border_f[i] * exposure_rat;
  for (int y = 1; y < patch_size_ + 1; ++y)
    for (int x = 0; x < patch_size_; ++x) patch_f[(y - 1) * patch_size_ + x] = patch_with_border_f[y * (patch_size_ + 2) + 1 + x];

  const V3 f{s.f[0], s.f[1], s.f[2]};
  // pClose = pClose / pClose[2]: Eigen's vector / scalar is an element-wise division
  V3 pClose;
  {
    const V3 pc = T_cur_ref.apply(f * min_idepth);
    pClose = {pc.x / pc.z, pc.y / pc.z, pc.z / pc.z};
  }
  V3 pFar = T_cur_ref.apply(f * max_idepth);
  if (pFar.z < 0.001 || max_idepth < min_idepth) return -1;
  pFar = {pFar.x / pFar.z, pFar.y / pFar.z, pFar.z / pFar.z};
  if (std::isnan((float)(pFar.x + pClose.x))) return -1;

  double px_close[2], px_far[2];
  {
    const double uc[3] = {pClose.x, pClose.y, 1.0}, uf[3] = {pFar.x, pFar.y, 1.0};
    orc_world2cam(cam, uc, px_close);  // world2cam(Vector2d) == world2cam(Vector3d(x, y, 1)): the divisions by 1 are exact
    orc_world2cam(cam, uf, px_far);
  }
  o.epl_start[0] = (int)px_close[0]; o.epl_start[1] = (int)px_close[1];
  px_close[0] = px_close[0] / (1 << search_level); px_close[1] = px_close[1] / (1 << search_level);
  o.epl_end[0] = (int)px_far[0]; o.epl_end[1] = (int)px_far[1];
  px_far[0] = px_far[0] / (1 << search_level); px_far[1] = px_far[1] / (1 << search_level);

  double incx = px_close[0] - px_far[0];
  double incy = px_close[1] - px_far[1];
  const double eplLength = std::sqrt(incx * incx + incy * incy);
  if ((!eplLength) > 0 || std::isinf(eplLength)) return -1;  // `!eplLength > 0` as written: true only for eplLength == 0
  if (eplLength > 100.0) {
    px_close[0] = px_far[0] + incx * 100.0 / eplLength;
    px_close[1] = px_far[1] + incy * 100.0 / eplLength;
  }
  incx *= 1.0 / eplLength;
  incy *= 1.0 / eplLength;
  px_far[0] -= incx; px_far[1] -= incy;
  px_close[0] += incx; px_close[1] += incy;
  if (eplLength < 2.0) {
    const double pad = (2.0 - (eplLength)) / 2.0f;
    px_far[0] -= incx * pad; px_far[1] -= incy * pad;
    px_close[0] += incx * pad; px_close[1] += incy * pad;
  }
  if (s.ftr_type == 2 || s.ftr_type == 1) {  // GRADIENT or EDGELET, options_.epi_search_edgelet_filtering = true
    double g0 = A[0] * s.grad[0] + A[1] * s.grad[1], g1 = A[2] * s.grad[0] + A[3] * s.grad[1];
    const double gn = std::sqrt(g0 * g0 + g1 * g1);
    g0 /= gn; g1 /= gn;
    double e0 = px_close[0] - px_far[0], e1 = px_close[1] - px_far[1];
    const double en = std::sqrt(e0 * e0 + e1 * e1);
    e0 /= en; e1 /= en;
    const double cosangle = std::fabs(g0 * e0 + g1 * e1);
    if (cosangle < 0.4) return -1;
  }

  double cpx = px_far[0], cpy = px_far[1];
  ZMNCC patchScore(patch_f);
  float zmncc_best = 0.1;
  float zmncc_second = zmncc_best;
  double uv_best[2] = {0, 0};
  float patch_cur[64];
  int loopCounter = 0;
  int loopCBest = -1, loopCSecond = -1;
  const Img cur{cur_levels[search_level], lw[search_level], lh[search_level]};
  while (((incx < 0) == (cpx > px_close[0]) && (incy < 0) == (cpy > px_close[1])) || loopCounter == 0) {
    const double px[2] = {cpx, cpy};
    if (!in_frame(cam, (int)px[0], (int)px[1], patch_size_, search_level)) {
      cpx += incx; cpy += incy; loopCounter++;
      continue;
    }
    create_patch(patch_cur, px, cur, halfpatch_size_);
    const float zmncc = patchScore.score(patch_cur);
    if (zmncc > zmncc_best) {
      zmncc_second = zmncc_best;
      uv_best[0] = px[0]; uv_best[1] = px[1];
      zmncc_best = zmncc;
      loopCSecond = loopCBest;
      loopCBest = loopCounter;
    } else if (zmncc > zmncc_second) {
      zmncc_second = zmncc;
      loopCSecond = loopCounter;
    }
    cpx += incx; cpy += incy; loopCounter++;
  }
  if (std::abs(loopCBest - loopCSecond) > 1.0f && 1.5f * zmncc_second > zmncc_best) return -4;

  if (zmncc_best > 0.8) {
    const double uv_best_0[2] = {uv_best[0] * (1 << search_level), uv_best[1] * (1 << search_level)};
    o.px_cur[0] = uv_best_0[0]; o.px_cur[1] = uv_best_0[1];
    double px_scaled[2] = {o.px_cur[0] / (1 << search_level), o.px_cur[1] / (1 << search_level)};
    double edir[2] = {px_close[0] - px_far[0], px_close[1] - px_far[1]};
    const double en = std::sqrt(edir[0] * edir[0] + edir[1] * edir[1]);
    edir[0] /= en; edir[1] /= en;
    bool result = klt_limited_1d(cur, patch_with_border_f, patch_f, align_max_iter, px_scaled, edir, NULL);
    float patch2D[64];
    std::memset(patch2D, 0, sizeof patch2D);  // uninitialised in the reference (see header)
    double dir_cur[2] = {A[0] * s.grad[0] + A[1] * s.grad[1], A[2] * s.grad[0] + A[3] * s.grad[1]};
    {
      const double n = std::sqrt(dir_cur[0] * dir_cur[0] + dir_cur[1] * dir_cur[1]);
      dir_cur[0] /= n; dir_cur[1] /= n;
    }
    if (!result) {
      double px_2d[2] = {o.px_cur[0] / (1 << search_level), o.px_cur[1] / (1 << search_level)};
      if (s.ftr_type != 1) {
        result = klt_limited_2d(cur, patch_with_border_f, patch_f, align_max_iter, px_2d, patch2D);
      } else {
        result = klt_limited_1d(cur, patch_with_border_f, patch_f, align_max_iter, px_2d, dir_cur, patch2D);
        if (result) result = orc_check_normal(cur_sobx[search_level], cur_soby[search_level], lw[search_level], px_2d, dir_cur, 0.7) != 0;
      }
      px_scaled[0] = px_2d[0]; px_scaled[1] = px_2d[1];
    } else {
      if (s.ftr_type != 1) {
        result = klt_limited_2d(cur, patch_with_border_f, patch_f, align_max_iter, px_scaled, patch2D);
      } else {
        result = klt_limited_1d(cur, patch_with_border_f, patch_f, align_max_iter, px_scaled, dir_cur, patch2D);
        if (result) result = orc_check_normal(cur_sobx[search_level], cur_soby[search_level], lw[search_level], px_scaled, dir_cur, 0.7) != 0;
      }
    }
    if (result) result = orc_check_ncc(patch_f, patch2D, 0.8) != 0;
    if (result) {
      o.px_cur[0] = px_scaled[0] * (1 << search_level); o.px_cur[1] = px_scaled[1] * (1 << search_level);
      double fc[3];
      orc_cam2world(cam, o.px_cur[0], o.px_cur[1], fc);
      if (depth_from_triangulation(T_cur_ref, f, V3{fc[0], fc[1], fc[2]}, result_depth)) return 1;
      return -2;
    }
    return -3;
  }
  return -4;
}

// DepthFilter::computeTau — src/depth_filter.cpp:539-555
double compute_tau(const SE3& T_ref_cur, const V3& f, double z, double px_error_angle) {
  const V3 t = T_ref_cur.t;
  const V3 a = f * z - t;
  const double t_norm = t.norm();
  const double a_norm = a.norm();
  const double alpha = std::acos(f.dot(t) / t_norm);
  const double beta = std::acos(a.dot(t * -1.0) / (t_norm * a_norm));
  const double beta_plus = beta + px_error_angle;
  const double gamma_plus = 3.14159265358979323846 - alpha - beta_plus;
  const double z_plus = t_norm * std::sin(beta_plus) / std::sin(gamma_plus);
  return (z_plus - z);
}

}  // namespace

extern "C" {

// DepthFilter::observeDepthRow — src/depth_filter.cpp:580-675 (the per-seed body; the threadReducer split does not change per-seed results)
void orc_depth_observe(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, double px_error_angle, int S,
                       const orc_seed_obs* seeds, int max_search_level, int align_max_iter, const uint8_t* const* const* ref_levels,
                       const uint8_t* const* cur_levels, const int* lw, const int* lh, const int16_t* const* cur_sobx, const int16_t* const* cur_soby,
                       orc_seed_result* out) {
  const SE3 Tc = SE3::from_rt(T_cur_w);
  std::vector<SE3> Tk;
  for (int k = 0; k < n_poses; ++k) Tk.push_back(SE3::from_rt(T_f_w + 12 * k));
  for (int i = 0; i < S; ++i) {
    const orc_seed_obs& s = seeds[i];
    orc_seed_result& r = out[i];
    std::memset(&r, 0, sizeof r);
    r.mu = s.mu; r.sigma2 = s.sigma2; r.is_valid = 1;
    const SE3 T_ref_cur = Tk[s.ref_pose].mul(Tc.inverse());
    const V3 f{s.f[0], s.f[1], s.f[2]};
    const V3 xyz_f = T_ref_cur.inverse().apply(f * (1.0 / s.mu));
    if (xyz_f.z < 0.0) continue;  // behind the camera
    {
      const double p[3] = {xyz_f.x, xyz_f.y, xyz_f.z};
      double px[2];
      orc_world2cam(cam, p, px);
      const int ox = (int)px[0], oy = (int)px[1];
      if (!(ox >= 0 && ox < cam->width && oy >= 0 && oy < cam->height)) continue;  // isInFrame(f2c(xyz_f).cast<int>())
    }
    r.is_update = 1;
    const float z_inv_min = s.mu + 2 * std::sqrt(s.sigma2);
    const float z_inv_max = std::max(s.mu - 2 * std::sqrt(s.sigma2), 0.00000001f);
    if (std::isnan(z_inv_min)) r.is_valid = 0;
    double z = 0;
    LineStereoOut o;
    std::memset(&o, 0, sizeof o);
    const SE3 T_cur_ref = Tc.mul(Tk[s.ref_pose].inverse());
    const int res = do_line_stereo(cam, T_cur_ref, s, ref_levels[s.ref_frame], cur_levels, lw, lh, cur_sobx, cur_soby, max_search_level, align_max_iter,
                                   1.0 / z_inv_min, 1.0 / s.mu, 1.0 / z_inv_max, z, o);
    r.res = res;
    r.search_level = o.search_level;
    if (res != 1) continue;  // it->b++, eplStart = eplEnd = (0,0): applied by the caller from res
    r.epl_start[0] = o.epl_start[0]; r.epl_start[1] = o.epl_start[1];
    r.epl_end[0] = o.epl_end[0]; r.epl_end[1] = o.epl_end[1];
    r.px_cur[0] = o.px_cur[0]; r.px_cur[1] = o.px_cur[1];
    r.z = z;
    const double tau = compute_tau(T_ref_cur, f, z, px_error_angle);
    const double tau_inverse = 0.5 * (1.0 / std::max(0.0000001, z - tau) - 1.0 / (z + tau));
    // DepthFilter::updateSeed(const float x, const float tau2, Seed*) — :528-537
    {
      const float x = 1. / z, tau2 = tau_inverse * tau_inverse;
      float id_var = r.sigma2 * 1.01f;
      const float w = tau2 / (tau2 + id_var);
      const float new_idepth = (1 - w) * x + w * r.mu;
      r.mu = (new_idepth < 0 ? (new_idepth > -1e-10 ? -1e-10 : new_idepth) : (new_idepth < 1e-10 ? 1e-10 : new_idepth));  // UNZERO
      id_var *= w;
      if (id_var < r.sigma2) r.sigma2 = id_var;
    }
  }
}

}  // extern "C"
