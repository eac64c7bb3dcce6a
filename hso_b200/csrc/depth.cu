// Row N3 on sm_100a: DepthFilter::observeDepthRow (src/depth_filter.cpp:580-675) for all seeds of a frame in one launch — visibility
// test, inverse-depth interval, Matcher::doLineStereo (src/matcher.cpp:802-1049): warp::getWarpMatrixAffine / getBestSearchLevel /
// warpAffine, exposure scaling, the epipolar segment at the search level, the ZMNCC_F scan (include/hso/vikit/patch_score.h:268-305) over
// warp::createPatch samples (src/matcher.cpp:159-196), the ambiguity test, KLTLimited1D / KLTLimited2D refinement (:1296-1606), checkNormal,
// checkNCC, depthFromTriangulation (:242-255) — then DepthFilter::computeTau (:539-555) and DepthFilter::updateSeed (:528-537).
//
// One warp per seed (the reference runs the seed list on 4 host threads, one seed at a time). Lane l owns pixels l and l+32 of every 8x8
// patch: a scan step is 8 byte gathers + 4 butterfly reductions, a KLT iteration 8 gathers + 3-4 reductions; butterfly (xor-shuffle) sums
// give every lane bit-identical totals, so the scalar control (best / second-best bookkeeping, damping, convergence) runs redundantly in all
// lanes with no broadcast and no divergence. The fp64 geometry is evaluated redundantly per lane as well (uniform, same latency as one lane).
// Float pipeline like the reference with a different summation order inside a patch (tolerances in tests/test_gpu_depth.py).
#include "hso_internal.h"

namespace hso {

constexpr int DEPTH_WARPS = 4;

namespace {

struct LevelImg { const uint8_t* data; int cols, rows; };

HSO_DEV float bilinear_wbr_prod(const uint8_t* it, int cols, float wTL, float wTR, float wBL, float wBR) {
  return wTL * __ldg(it) + wTR * __ldg(it + 1) + wBL * __ldg(it + cols) + wBR * __ldg(it + cols + 1);
}

// Matcher::KLTLimited1D (src/matcher.cpp:1454-1606). tp: the lane's two pixels of targetPatch (nullptr = NULL in the reference).
HSO_DEV bool klt_limited_1d(int lane, const LevelImg& img, const float* patch /*10x10*/, const float* refv, int n_iter, double* px, double d0, double d1,
                            float* tp) {
  float dv[2], wgt[2];
  float h00 = 0, h01 = 0, h11 = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int p = lane + 32 * k;
    const float* it = patch + ((p >> 3) + 1) * 10 + 1 + (p & 7);
    dv[k] = (float)(0.5 * (d0 * (double)(it[1] - it[-1]) + d1 * (double)(it[10] - it[-10])));
    wgt[k] = sqrtf((float)(250.0 / (250.0 + (double)(dv[k] * dv[k]))));
    h00 += (dv[k] * dv[k]) * wgt[k]; h01 += dv[k] * wgt[k]; h11 += wgt[k];
  }
  h00 = warp_sum(h00); h01 = warp_sum(h01); h11 = warp_sum(h11);
  const float H0 = (float)((double)h00 * (1 + 0.001)), H3 = (float)((double)h11 * (1 + 0.001));
  const float det = H0 * H3 - h01 * h01;
  const float id = 1.0f / det;
  const float Hi0 = H3 * id, Hi1 = -h01 * id, Hi2 = -h01 * id, Hi3 = H0 * id;
  float mean_diff = 0;
  float bestU = (float)px[0], bestV = (float)px[1];
  float bestEnergy = 1e8f;
  float sb0 = 0, sb1 = 0;
  float uBak = bestU, vBak = bestV, meanBak = mean_diff;
  bool nan_fail = false;
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = __double2int_rd((double)bestU), v_r = __double2int_rd((double)bestV);
    if (u_r < 4 || v_r < 4 || u_r >= img.cols - 4 || v_r >= img.rows - 4) break;
    if (isnan(bestU) || isnan(bestV)) { nan_fail = true; break; }
    const float sx = bestU - (float)u_r, sy = bestV - (float)v_r;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy)), wTR = (float)(sx * (1.0 - sy)), wBL = (float)((1.0 - sx) * sy), wBR = sx * sy;
    float j0 = 0, j1 = 0, energy = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = lane + 32 * k;
      const uint8_t* it = img.data + (v_r + (p >> 3) - 4) * img.cols + u_r - 4 + (p & 7);
      const float sp = bilinear_wbr_prod(it, img.cols, wTL, wTR, wBL, wBR);
      const float res = sp - refv[k] + mean_diff;
      j0 -= res * dv[k] * wgt[k];
      j1 -= res * wgt[k];
      energy += res * res * wgt[k];
      if (tp) tp[k] = sp;
    }
    j0 = warp_sum(j0); j1 = warp_sum(j1); energy = warp_sum(energy);
    if (energy > bestEnergy) {
      sb0 *= 0.5f; sb1 *= 0.5f;
      bestU = (float)((double)uBak + (double)sb0 * d0);
      bestV = (float)((double)vBak + (double)sb0 * d1);
      mean_diff = meanBak + sb1;
    } else {
      float s0 = Hi0 * j0 + Hi1 * j1, s1 = Hi2 * j0 + Hi3 * j1;
      if (s0 < -0.5f) s0 = -0.5f; else if (s0 > 0.5f) s0 = 0.5f;
      if (!isfinite(s0)) { s0 = 0; s1 = 0; }
      uBak = bestU; vBak = bestV; meanBak = mean_diff;
      sb0 = s0; sb1 = s1;
      bestU = (float)((double)bestU + (double)s0 * d0);
      bestV = (float)((double)bestV + (double)s0 * d1);
      mean_diff += s1;
      bestEnergy = energy;
    }
    if (fabsf(sb0) < 0.01f) break;
  }
  if (nan_fail) return false;  // `return false` before targetPxEstimate is written
  px[0] = (double)bestU; px[1] = (double)bestV;
  return !(bestEnergy > 650.f * 64.f);
}

// Matcher::KLTLimited2D (src/matcher.cpp:1296-1450)
HSO_DEV bool klt_limited_2d(int lane, const LevelImg& img, const float* patch, const float* refv, int n_iter, double* px, float* tp) {
  float dx[2], dy[2], wgt[2];
  float h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int p = lane + 32 * k;
    const float* it = patch + ((p >> 3) + 1) * 10 + 1 + (p & 7);
    dx[k] = (float)(0.5 * (double)(it[1] - it[-1]));
    dy[k] = (float)(0.5 * (double)(it[10] - it[-10]));
    wgt[k] = sqrtf((float)(250.0 / (250.0 + (double)(dx[k] * dx[k] + dy[k] * dy[k]))));
    h00 += (dx[k] * dx[k]) * wgt[k]; h01 += (dx[k] * dy[k]) * wgt[k]; h02 += dx[k] * wgt[k];
    h11 += (dy[k] * dy[k]) * wgt[k]; h12 += dy[k] * wgt[k]; h22 += wgt[k];
  }
  h00 = warp_sum(h00); h01 = warp_sum(h01); h02 = warp_sum(h02); h11 = warp_sum(h11); h12 = warp_sum(h12); h22 = warp_sum(h22);
  const float H[9] = {(float)((double)h00 * (1 + 0.001)), h01, h02, h01, (float)((double)h11 * (1 + 0.001)), h12, h02, h12, (float)((double)h22 * (1 + 0.001))};
  float Hi[9];
  {  // Eigen fixed-size 3x3 inverse: cofactors / determinant
    const float c00 = H[4] * H[8] - H[5] * H[7], c10 = H[5] * H[6] - H[3] * H[8], c20 = H[3] * H[7] - H[4] * H[6];
    const float det = H[0] * c00 + H[1] * c10 + H[2] * c20;
    const float id = 1.0f / det;
    Hi[0] = c00 * id; Hi[1] = (H[2] * H[7] - H[1] * H[8]) * id; Hi[2] = (H[1] * H[5] - H[2] * H[4]) * id;
    Hi[3] = c10 * id; Hi[4] = (H[0] * H[8] - H[2] * H[6]) * id; Hi[5] = (H[2] * H[3] - H[0] * H[5]) * id;
    Hi[6] = c20 * id; Hi[7] = (H[1] * H[6] - H[0] * H[7]) * id; Hi[8] = (H[0] * H[4] - H[1] * H[3]) * id;
  }
  float mean_diff = 0;
  float bestU = (float)px[0], bestV = (float)px[1];
  float bestEnergy = 1e8f;
  float sb0 = 0, sb1 = 0, sb2 = 0;
  float uBak = bestU, vBak = bestV, meanBak = mean_diff;
  bool nan_fail = false;
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = __double2int_rd((double)bestU), v_r = __double2int_rd((double)bestV);
    if (u_r < 4 || v_r < 4 || u_r >= img.cols - 4 || v_r >= img.rows - 4) break;
    if (isnan(bestU) || isnan(bestV)) { nan_fail = true; break; }
    const float sx = bestU - (float)u_r, sy = bestV - (float)v_r;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy)), wTR = (float)(sx * (1.0 - sy)), wBL = (float)((1.0 - sx) * sy), wBR = sx * sy;
    float j0 = 0, j1 = 0, j2 = 0, energy = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = lane + 32 * k;
      const uint8_t* it = img.data + (v_r + (p >> 3) - 4) * img.cols + u_r - 4 + (p & 7);
      const float sp = bilinear_wbr_prod(it, img.cols, wTL, wTR, wBL, wBR);
      const float res = sp - refv[k] + mean_diff;
      j0 -= res * dx[k] * wgt[k];
      j1 -= res * dy[k] * wgt[k];
      j2 -= res * wgt[k];
      energy += res * res * wgt[k];
      tp[k] = sp;
    }
    j0 = warp_sum(j0); j1 = warp_sum(j1); j2 = warp_sum(j2); energy = warp_sum(energy);
    if (energy > bestEnergy) {
      sb0 *= 0.5f; sb1 *= 0.5f; sb2 *= 0.5f;
      bestU = uBak + sb0; bestV = vBak + sb1; mean_diff = meanBak + sb2;
    } else {
      float s0 = Hi[0] * j0 + Hi[1] * j1 + Hi[2] * j2, s1 = Hi[3] * j0 + Hi[4] * j1 + Hi[5] * j2, s2 = Hi[6] * j0 + Hi[7] * j1 + Hi[8] * j2;
      if (s0 < -0.5f) s0 = -0.5f; else if (s0 > 0.5f) s0 = 0.5f;
      if (s1 < -0.5f) s1 = -0.5f; else if (s1 > 0.5f) s1 = 0.5f;
      if (!isfinite(s0)) { s0 = 0; s1 = 0; s2 = 0; }
      uBak = bestU; vBak = bestV; meanBak = mean_diff;
      sb0 = s0; sb1 = s1; sb2 = s2;
      bestU += s0; bestV += s1; mean_diff += s2;
      bestEnergy = energy;
    }
    if ((double)(sb0 * sb1) < 0.01 * 0.01) break;  // the reference tests the PRODUCT of the two components (:1437)
  }
  if (nan_fail) return false;
  px[0] = (double)bestU; px[1] = (double)bestV;
  return !(bestEnergy > 650.f * 64.f);
}

}  // namespace

__global__ void __launch_bounds__(DEPTH_WARPS * 32) k_depth_observe(const DepthKParams P, const hso_seed_obs* __restrict__ seeds,
                                                                    const uint8_t* const* __restrict__ ref_pyr, hso_seed_result* __restrict__ out) {
  __shared__ float s_patch[DEPTH_WARPS][100];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * DEPTH_WARPS + warp;
  if (m >= P.S) return;
  const hso_seed_obs s = seeds[m];
  float* patch = s_patch[warp];
  hso_seed_result r;
  r.is_update = 0; r.is_valid = 1; r.res = 0; r.search_level = 0;
  r.epl_start[0] = r.epl_start[1] = r.epl_end[0] = r.epl_end[1] = 0;
  r.mu = s.mu; r.sigma2 = s.sigma2; r.z = 0; r.px_cur[0] = r.px_cur[1] = 0;

  // ---- DepthFilter::observeDepthRow head (src/depth_filter.cpp:591-624) ---------------------------------------------------------
  const Se3d Tc = se3_from_rt(P.T_cur_w);
  const Se3d Tr = se3_from_rt(P.T_f_w + 12 * s.ref_pose);
  const Se3d T_ref_cur = se3_mul(Tr, se3_inverse(Tc));
  bool visible;
  {
    const Se3d Ti = se3_inverse(T_ref_cur);
    const double inv = 1.0 / (double)s.mu;
    double X, Y, Z;
    se3_apply(Ti, inv * s.f[0], inv * s.f[1], inv * s.f[2], X, Y, Z);
    visible = !(Z < 0.0);
    if (visible) {
      double pu, pv;
      world2cam_exact(P.cam, X, Y, Z, pu, pv);
      const int ox = (int)pu, oy = (int)pv;
      visible = ox >= 0 && ox < P.cam.width && oy >= 0 && oy < P.cam.height;
    }
  }
  if (!visible) {
    if (lane == 0) out[m] = r;
    return;
  }
  r.is_update = 1;
  const float sq = sqrtf(s.sigma2);
  const float z_inv_min = s.mu + 2.f * sq;
  const float z_inv_max = fmaxf(s.mu - 2.f * sq, 0.00000001f);
  if (isnan(z_inv_min)) r.is_valid = 0;
  const double min_d = 1.0 / (double)z_inv_min, prior_d = 1.0 / (double)s.mu, max_d = 1.0 / (double)z_inv_max;

  // ---- Matcher::doLineStereo (src/matcher.cpp:802-1049) ------------------------------------------------------------------------------
  const Se3d Tcr = se3_mul(Tc, se3_inverse(Tr));
  double A[4];
  int sl = 0;
  {  // warp::getWarpMatrixAffine (:46-72) at the prior depth, getBestSearchLevel (:74-85)
    const int halfpatch = 5;
    const double xr = s.f[0] * prior_d, yr = s.f[1] * prior_d, zr = s.f[2] * prior_d;
    const int ratio = 1 << s.level;
    double dux, duy, duz, dvx, dvy, dvz;
    cam2world(P.cam, s.px[0] + (double)(halfpatch * ratio), s.px[1], dux, duy, duz);
    cam2world(P.cam, s.px[0], s.px[1] + (double)(halfpatch * ratio), dvx, dvy, dvz);
    const double sdu = zr / duz, sdv = zr / dvz;
    double ax, ay, az, bx, by, bz, cx, cy, cz;
    se3_apply(Tcr, xr, yr, zr, ax, ay, az);
    se3_apply(Tcr, dux * sdu, duy * sdu, duz * sdu, bx, by, bz);
    se3_apply(Tcr, dvx * sdv, dvy * sdv, dvz * sdv, cx, cy, cz);
    double pcu, pcv, puu, puv, pvu, pvv;
    world2cam_exact(P.cam, ax, ay, az, pcu, pcv);
    world2cam_exact(P.cam, bx, by, bz, puu, puv);
    world2cam_exact(P.cam, cx, cy, cz, pvu, pvv);
    A[0] = (puu - pcu) / halfpatch; A[2] = (puv - pcv) / halfpatch;
    A[1] = (pvu - pcu) / halfpatch; A[3] = (pvv - pcv) / halfpatch;
    double D = A[0] * A[3] - A[1] * A[2];
    while (D > 3.0 && sl < P.max_search_level) { sl += 1; D *= 0.25; }
  }
  r.search_level = sl;
  {  // warp::warpAffine float overload (:120-155), halfpatch 5 => 10x10, then exposure scaling (:820-830)
    const double invdet = 1.0 / (A[0] * A[3] - A[1] * A[2]);
    const float i00 = (float)(A[3] * invdet), i01 = (float)(-A[1] * invdet), i10 = (float)(-A[2] * invdet), i11 = (float)(A[0] * invdet);
    const bool bad = isnan(i00);
    const int rl = s.level;
    const float pxr0 = (float)(s.px[0] / (double)(1 << rl)), pxr1 = (float)(s.px[1] / (double)(1 << rl));
    const float scale_target = (float)(1 << sl);
    const uint8_t* ref = ref_pyr[m] + P.g.off[rl];
    const int cols = P.g.w[rl], rows = P.g.h[rl];
    const bool scale_patch = fabsf(s.exposure_rat * 128.f - 128.f) > 30.0f;
    for (int idx = lane; idx < 100; idx += 32) {
      const int y = idx / 10, x = idx - y * 10;
      const float p0 = (float)(x - 5) * scale_target, p1 = (float)(y - 5) * scale_target;
      const float q0 = (i00 * p0 + i01 * p1) + pxr0;
      const float q1 = (i10 * p0 + i11 * p1) + pxr1;
      float val = 0.f;
      if (!bad && !(q0 < 0 || q1 < 0 || q0 >= cols - 1 || q1 >= rows - 1)) {
        // interpolateMat_8u (include/hso/vikit/vision.h:49-65): w11 = 1 - w00 - w01 - w10
        const float xf = floorf(q0), yf = floorf(q1);
        const int xi = (int)xf, yi = (int)yf;
        const float sx = q0 - xf, sy = q1 - yf;
        const float w00 = (1.0f - sx) * (1.0f - sy), w01 = (1.0f - sx) * sy, w10 = sx * (1.0f - sy);
        const float w11 = 1.0f - w00 - w01 - w10;
        const uint8_t* p = ref + yi * cols + xi;
        val = w00 * __ldg(p) + w01 * __ldg(p + cols) + w10 * __ldg(p + 1) + w11 * __ldg(p + cols + 1);
      }
      if (scale_patch) val = val * s.exposure_rat;
      patch[idx] = val;
    }
  }
  __syncwarp();
  float refv[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int p = lane + 32 * k;
    refv[k] = patch[((p >> 3) + 1) * 10 + 1 + (p & 7)];
  }

  int res = 0;
  double px_close0 = 0, px_close1 = 0, px_far0 = 0, px_far1 = 0, incx = 0, incy = 0;
  int es0 = 0, es1 = 0, ee0 = 0, ee1 = 0;
  {  // the epipolar segment (:836-905)
    double cx_, cy_, cz_, fx_, fy_, fz_;
    se3_apply(Tcr, s.f[0] * min_d, s.f[1] * min_d, s.f[2] * min_d, cx_, cy_, cz_);
    cx_ = cx_ / cz_; cy_ = cy_ / cz_;
    se3_apply(Tcr, s.f[0] * max_d, s.f[1] * max_d, s.f[2] * max_d, fx_, fy_, fz_);
    if (fz_ < 0.001 || max_d < min_d) res = -1;
    if (res == 0) {
      fx_ = fx_ / fz_; fy_ = fy_ / fz_;
      if (isnan((float)(fx_ + cx_))) res = -1;
    }
    if (res == 0) {
      world2cam_exact(P.cam, cx_, cy_, 1.0, px_close0, px_close1);
      world2cam_exact(P.cam, fx_, fy_, 1.0, px_far0, px_far1);
      es0 = (int)px_close0; es1 = (int)px_close1;
      ee0 = (int)px_far0; ee1 = (int)px_far1;
      const double lev = (double)(1 << sl);
      px_close0 /= lev; px_close1 /= lev; px_far0 /= lev; px_far1 /= lev;
      incx = px_close0 - px_far0; incy = px_close1 - px_far1;
      const double eplLength = sqrt(incx * incx + incy * incy);
      if (eplLength == 0.0 || isinf(eplLength)) res = -1;  // `!eplLength > 0 || isinf` as written (:870)
      if (res == 0) {
        if (eplLength > 100.0) {
          px_close0 = px_far0 + incx * 100.0 / eplLength;
          px_close1 = px_far1 + incy * 100.0 / eplLength;
        }
        incx *= 1.0 / eplLength; incy *= 1.0 / eplLength;
        px_far0 -= incx; px_far1 -= incy; px_close0 += incx; px_close1 += incy;
        if (eplLength < 2.0) {
          const double pad = (2.0 - eplLength) / 2.0;
          px_far0 -= incx * pad; px_far1 -= incy * pad; px_close0 += incx * pad; px_close1 += incy * pad;
        }
        if (s.ftr_type == 2 || s.ftr_type == 1) {  // epi_search_edgelet_filtering (:908-914)
          double g0 = A[0] * s.grad[0] + A[1] * s.grad[1], g1 = A[2] * s.grad[0] + A[3] * s.grad[1];
          const double gn = sqrt(g0 * g0 + g1 * g1);
          g0 /= gn; g1 /= gn;
          double e0 = px_close0 - px_far0, e1 = px_close1 - px_far1;
          const double en = sqrt(e0 * e0 + e1 * e1);
          e0 /= en; e1 /= en;
          if (fabs(g0 * e0 + g1 * e1) < 0.4) res = -1;
        }
      }
    }
  }
  const LevelImg cur{P.cur_pyr + P.g.off[sl], P.g.w[sl], P.g.h[sl]};
  float zmncc_best = 0.1f, zmncc_second = 0.1f;
  double uv_best0 = 0, uv_best1 = 0;
  if (res == 0) {
    // ---- ZMNCC scan along the segment (:917-968) -------------------------------------------------------------------------------------
    float hostMean = warp_sum(refv[0] + refv[1]) / 64.f;
    const float h0 = refv[0] - hostMean, h1 = refv[1] - hostMean;
    const float den1 = warp_sum(h0 * h0 + h1 * h1);
    double cpx = px_far0, cpy = px_far1;
    int loopCounter = 0, loopCBest = -1, loopCSecond = -1;
    const int lw_ = P.cam.width / (1 << sl), lh_ = P.cam.height / (1 << sl);
    while ((((incx < 0) == (cpx > px_close0)) && ((incy < 0) == (cpy > px_close1))) || loopCounter == 0) {
      const int ox = (int)cpx, oy = (int)cpy;
      if (ox >= 8 && ox < lw_ - 8 && oy >= 8 && oy < lh_ - 8) {  // isInFrame(px.cast<int>(), patch_size_, search_level_)
        // warp::createPatch (:159-196): w_br = 1 - tl - tr - bl
        const float u_cur = (float)cpx, v_cur = (float)cpy;
        const int ui = (int)floorf(u_cur), vi = (int)floorf(v_cur);
        const float su = u_cur - (float)ui, sv = v_cur - (float)vi;
        const float wtl = (float)((1.0 - su) * (1.0 - sv)), wtr = (float)(su * (1.0 - sv)), wbl = (float)((1.0 - su) * sv);
        const float wbr = (float)(1.0 - wtl - wtr - wbl);
        float t[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int p = lane + 32 * k;
          const uint8_t* it = cur.data + (vi - 4 + (p >> 3)) * cur.cols + (ui - 4) + (p & 7);
          t[k] = bilinear_wbr_prod(it, cur.cols, wtl, wtr, wbl, wbr);
        }
        const float tmean = warp_sum(t[0] + t[1]) / 64.f;
        const float t0 = t[0] - tmean, t1 = t[1] - tmean;
        const float num = warp_sum(h0 * t0 + h1 * t1);
        const float den2 = warp_sum(t0 * t0 + t1 * t1);
        const float zmncc = (float)((double)num / ((double)sqrtf(den1 * den2) + 1e-12));
        if (zmncc > zmncc_best) {
          zmncc_second = zmncc_best;
          uv_best0 = cpx; uv_best1 = cpy;
          zmncc_best = zmncc;
          loopCSecond = loopCBest;
          loopCBest = loopCounter;
        } else if (zmncc > zmncc_second) {
          zmncc_second = zmncc;
          loopCSecond = loopCounter;
        }
      }
      cpx += incx; cpy += incy; loopCounter++;
    }
    if (abs(loopCBest - loopCSecond) > 1 && 1.5f * zmncc_second > zmncc_best) res = -4;
  }
  double z = 0, pxc0 = 0, pxc1 = 0;
  if (res == 0) {
    if ((double)zmncc_best > 0.8) {  // float against the double literal, as written (:975)
      // ---- refinement (:977-1047) ----------------------------------------------------------------------------------------------------
      const double lev = (double)(1 << sl);
      pxc0 = uv_best0 * lev; pxc1 = uv_best1 * lev;
      double pxs[2] = {pxc0 / lev, pxc1 / lev};
      double e0 = px_close0 - px_far0, e1 = px_close1 - px_far1;
      const double en = sqrt(e0 * e0 + e1 * e1);
      e0 /= en; e1 /= en;
      bool result = klt_limited_1d(lane, cur, patch, refv, P.max_iter, pxs, e0, e1, nullptr);
      float tp[2] = {0.f, 0.f};  // patch2D: uninitialised in the reference when no iteration runs; defined as 0 here
      double dc0 = A[0] * s.grad[0] + A[1] * s.grad[1], dc1 = A[2] * s.grad[0] + A[3] * s.grad[1];
      {
        const double n = sqrt(dc0 * dc0 + dc1 * dc1);
        dc0 /= n; dc1 /= n;
      }
      double p2[2] = {pxc0 / lev, pxc1 / lev};
      double* tgt = result ? pxs : p2;  // !result: restart from the scan position (:990), else continue from the 1-D result (:1008)
      if (s.ftr_type != 1) {
        result = klt_limited_2d(lane, cur, patch, refv, P.max_iter, tgt, tp);
      } else {
        result = klt_limited_1d(lane, cur, patch, refv, P.max_iter, tgt, dc0, dc1, tp);
        if (result) {  // checkNormal(cur_frame, search_level_, px, dir_cur, 0.7) (:406-440)
          const float uf = (float)tgt[0], vf = (float)tgt[1];
          const int ui = __float2int_rd(uf), vi = __float2int_rd(vf);
          const float sx = uf - (float)ui, sy = vf - (float)vi;
          const float wTL = (float)((1.0 - sx) * (1.0 - sy)), wTR = (float)(sx * (1.0 - sy)), wBL = (float)((1.0 - sx) * sy);
          const float wBR = (float)(1.0 - wTL - wTR - wBL);
          const int16_t* gxp = P.cur_sobel + P.sobel_off[sl];
          const int16_t* gyp = gxp + (size_t)cur.cols * cur.rows;
          const size_t o = (size_t)vi * cur.cols + ui;
          const double nx = (double)wTL * gxp[o] + (double)wTR * gxp[o + 1] + (double)wBL * gxp[o + cur.cols] + (double)wBR * gxp[o + cur.cols + 1];
          const double ny = (double)wTL * gyp[o] + (double)wTR * gyp[o + 1] + (double)wBL * gyp[o + cur.cols] + (double)wBR * gyp[o + cur.cols + 1];
          const double nn = sqrt(nx * nx + ny * ny);
          result = (dc0 * (nx / nn) + dc1 * (ny / nn)) > (double)0.7f;
        }
      }
      pxs[0] = tgt[0]; pxs[1] = tgt[1];
      if (result) {  // checkNCC(patch_f_, patch2D, 0.8) (:379-404)
        float mean1 = warp_sum(refv[0] + refv[1]), mean2 = warp_sum(tp[0] + tp[1]);
        mean1 /= 64.f; mean2 /= 64.f;
        float num = 0, d1 = 0, d2 = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float a1 = refv[k] - mean1, a2 = tp[k] - mean2;
          num += a1 * a2; d1 += a1 * a1; d2 += a2 * a2;
        }
        num = warp_sum(num); d1 = warp_sum(d1); d2 = warp_sum(d2);
        result = ((double)num / ((double)sqrtf(d1 * d2) + 1e-12)) > (double)0.8f;
      }
      if (result) {
        pxc0 = pxs[0] * lev; pxc1 = pxs[1] * lev;
        // depthFromTriangulation(T_cur_ref, ref_ftr.f, cam2world(px_cur_), depth) (:242-255)
        double fcx, fcy, fcz;
        cam2world(P.cam, pxc0, pxc1, fcx, fcy, fcz);
        double a0x, a0y, a0z;
        {
          double R[9];
          quat_to_R(Tcr.q, R);
          a0x = R[0] * s.f[0] + R[1] * s.f[1] + R[2] * s.f[2];
          a0y = R[3] * s.f[0] + R[4] * s.f[1] + R[5] * s.f[2];
          a0z = R[6] * s.f[0] + R[7] * s.f[1] + R[8] * s.f[2];
        }
        const double m00 = a0x * a0x + a0y * a0y + a0z * a0z, m01 = a0x * fcx + a0y * fcy + a0z * fcz, m11 = fcx * fcx + fcy * fcy + fcz * fcz;
        const double det = m00 * m11 - m01 * m01;
        if (det < 0.000001) {
          res = -2;
        } else {
          const double b0 = a0x * Tcr.tx + a0y * Tcr.ty + a0z * Tcr.tz, b1 = fcx * Tcr.tx + fcy * Tcr.ty + fcz * Tcr.tz;
          z = fabs(-((m11 / det) * b0 + (-m01 / det) * b1));
          res = 1;
        }
      } else {
        res = -3;
      }
    } else {
      res = -4;
    }
  }
  r.res = res;
  if (res == 1) {
    r.epl_start[0] = es0; r.epl_start[1] = es1; r.epl_end[0] = ee0; r.epl_end[1] = ee1;
    r.px_cur[0] = pxc0; r.px_cur[1] = pxc1;
    r.z = z;
    // DepthFilter::computeTau (src/depth_filter.cpp:539-555)
    const double tx = T_ref_cur.tx, ty = T_ref_cur.ty, tz = T_ref_cur.tz;
    const double ax = s.f[0] * z - tx, ay = s.f[1] * z - ty, az = s.f[2] * z - tz;
    const double t_norm = sqrt(tx * tx + ty * ty + tz * tz), a_norm = sqrt(ax * ax + ay * ay + az * az);
    const double alpha = acos((s.f[0] * tx + s.f[1] * ty + s.f[2] * tz) / t_norm);
    const double beta = acos((ax * -tx + ay * -ty + az * -tz) / (t_norm * a_norm));
    const double beta_plus = beta + P.px_error_angle;
    const double gamma_plus = 3.14159265358979323846 - alpha - beta_plus;
    const double z_plus = t_norm * sin(beta_plus) / sin(gamma_plus);
    const double tau = z_plus - z;
    const double tau_inverse = 0.5 * (1.0 / fmax(0.0000001, z - tau) - 1.0 / (z + tau));
    // DepthFilter::updateSeed(const float x, const float tau2, Seed*) (:528-537)
    const float x = (float)(1. / z), tau2 = (float)(tau_inverse * tau_inverse);
    float id_var = r.sigma2 * 1.01f;
    const float w = tau2 / (tau2 + id_var);
    const float new_idepth = (1.f - w) * x + w * r.mu;
    const double nz = (double)new_idepth;
    r.mu = (float)(new_idepth < 0 ? (nz > -1e-10 ? -1e-10 : nz) : (nz < 1e-10 ? 1e-10 : nz));  // UNZERO
    id_var *= w;
    if (id_var < r.sigma2) r.sigma2 = id_var;
  }
  if (lane == 0) out[m] = r;
}

cudaError_t launch_depth_observe(const DepthKParams& p, const hso_seed_obs* seeds_dev, const uint8_t* const* ref_pyr_dev, hso_seed_result* out_dev,
                                 cudaStream_t stream, uint64_t* launches) {
  k_depth_observe<<<(p.S + DEPTH_WARPS - 1) / DEPTH_WARPS, DEPTH_WARPS * 32, 0, stream>>>(p, seeds_dev, ref_pyr_dev, out_dev);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace hso
