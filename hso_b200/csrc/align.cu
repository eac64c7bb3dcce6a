// Direct patch matching on sm_100a: the part of hso::Matcher::findMatchDirect after the host-side choice of the reference
// observation and affine warp (src/matcher.cpp:310-375): warp::warpAffine float overload (:120-155, interpolateMat_8u
// include/hso/vikit/vision.h:49-65), optional exposure scaling (:317-335), feature_alignment::align2D / align1D float overloads
// (src/feature_alignment.cpp:464-605 / :164-308), checkNormal (:406-440), checkNCC (:379-404), the 20-px displacement gate.
//
// One warp per candidate: lane l owns pixels l and l+32 of the 8x8 patch (rows l/8 and l/8+4). The 10x10 bordered reference
// patch lives in shared memory (400 B per warp); the inverse-compositional Hessian, the per-iteration J^T r and chi^2 and the
// NCC sums are butterfly (xor-shuffle) reductions, so every lane holds bit-identical sums and runs the 3x3 / 2x2 update
// redundantly — no broadcast, no divergence. The current-level window (<= ~11x11 px) is gathered from the device-resident
// pyramid through L1/L2. Float throughout like the reference; the summation order differs from its sequential loop
// (tolerance in tests/test_gpu_align.py).
#include "hso_internal.h"

namespace hso {

constexpr int ALIGN_WARPS = 8;

struct AlignKParams {
  PyrGeom g;
  const uint8_t* cur_pyr;
  const int16_t* cur_sobel;  // [level 0..2][gx | gy] or nullptr
  size_t sobel_off[3];
  int max_iter;
  int M;
};

// include/hso/vikit/vision.h:49-65
HSO_DEV float interpolate_8u(const uint8_t* data, int stride, float u, float v) {
  const float xf = floorf(u), yf = floorf(v);
  const int x = (int)xf, y = (int)yf;
  const float sx = u - xf, sy = v - yf;
  const float w00 = (1.0f - sx) * (1.0f - sy);
  const float w01 = (1.0f - sx) * sy;
  const float w10 = sx * (1.0f - sy);
  const float w11 = 1.0f - w00 - w01 - w10;
  const uint8_t* p = data + y * stride + x;
  return w00 * __ldg(p) + w01 * __ldg(p + stride) + w10 * __ldg(p + 1) + w11 * __ldg(p + stride + 1);
}

HSO_DEV void inv3(const float* H, float* Hi) {  // Eigen fixed-size 3x3 inverse: cofactors / determinant
  const float c00 = H[4] * H[8] - H[5] * H[7];
  const float c10 = H[5] * H[6] - H[3] * H[8];
  const float c20 = H[3] * H[7] - H[4] * H[6];
  const float det = H[0] * c00 + H[1] * c10 + H[2] * c20;
  const float id = 1.0f / det;
  Hi[0] = c00 * id; Hi[1] = (H[2] * H[7] - H[1] * H[8]) * id; Hi[2] = (H[1] * H[5] - H[2] * H[4]) * id;
  Hi[3] = c10 * id; Hi[4] = (H[0] * H[8] - H[2] * H[6]) * id; Hi[5] = (H[2] * H[3] - H[0] * H[5]) * id;
  Hi[6] = c20 * id; Hi[7] = (H[1] * H[6] - H[0] * H[7]) * id; Hi[8] = (H[0] * H[4] - H[1] * H[3]) * id;
}

__global__ void __launch_bounds__(ALIGN_WARPS * 32) k_align(const AlignKParams P, const AlignJobDev* __restrict__ jobs, hso_align_result* __restrict__ out) {
  __shared__ float s_patch[ALIGN_WARPS][100];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * ALIGN_WARPS + warp;
  if (m >= P.M) return;
  const AlignJobDev& jd = jobs[m];
  const hso_align_job& jb = jd.job;
  if (jb.ref_level < 0) {  // produced on the device by k_reproject: findMatchDirect returns false before the alignment (src/matcher.cpp:276-291)
    if (lane == 0) { out[m].ok = 0; out[m].align_converged = 0; out[m].px_cur[0] = jb.px_cur[0]; out[m].px_cur[1] = jb.px_cur[1]; out[m].h_inv = 0; }
    return;
  }
  float* patch = s_patch[warp];
  const int rl = jb.ref_level, sl = jb.search_level;

  // ---- warp::warpAffine, halfpatch 5 => 10x10 (src/matcher.cpp:120-155) ------------------------------------------------------
  {
    const double a0 = jb.A_cur_ref[0], a1 = jb.A_cur_ref[1], a2 = jb.A_cur_ref[2], a3 = jb.A_cur_ref[3];
    const double invdet = 1.0 / (a0 * a3 - a1 * a2);
    const float i00 = (float)(a3 * invdet), i01 = (float)(-a1 * invdet), i10 = (float)(-a2 * invdet), i11 = (float)(a0 * invdet);
    const bool bad = isnan(i00);  // the reference prints and leaves the patch unwritten; we zero it (NCC then fails)
    const float pxr0 = (float)(jb.px_ref[0] / (double)(1 << rl)), pxr1 = (float)(jb.px_ref[1] / (double)(1 << rl));
    const float scale_target = (float)(1 << sl);
    const uint8_t* ref = jd.ref_pyr + P.g.off[rl];
    const int cols = P.g.w[rl], rows = P.g.h[rl];
    for (int idx = lane; idx < 100; idx += 32) {
      const int y = idx / 10, x = idx - y * 10;
      const float p0 = (float)(x - 5) * scale_target, p1 = (float)(y - 5) * scale_target;
      const float q0 = (i00 * p0 + i01 * p1) + pxr0;
      const float q1 = (i10 * p0 + i11 * p1) + pxr1;
      float val = 0.f;
      if (!bad && !(q0 < 0 || q1 < 0 || q0 >= cols - 1 || q1 >= rows - 1)) val = interpolate_8u(ref, cols, q0, q1);
      if (jb.scale_patch) val = val * jb.exposure_rat;  // src/matcher.cpp:317-330
      patch[idx] = val;
    }
  }
  __syncwarp();

  // ---- template derivatives, weights, Hessian (feature_alignment.cpp:487-513 / :186-207) -------------------------------------
  const bool edgelet = jb.type == 1;
  float dirx = 0.f, diry = 0.f;
  double dir_d0 = 0, dir_d1 = 0;
  if (edgelet) {
    dir_d0 = jb.A_cur_ref[0] * jb.grad[0] + jb.A_cur_ref[1] * jb.grad[1];
    dir_d1 = jb.A_cur_ref[2] * jb.grad[0] + jb.A_cur_ref[3] * jb.grad[1];
    const double n = sqrt(dir_d0 * dir_d0 + dir_d1 * dir_d1);
    dir_d0 /= n; dir_d1 /= n;
    dirx = (float)dir_d0; diry = (float)dir_d1;
  }
  float refv[2], jx[2], jy[2], wgt[2];
  float h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int p = lane + 32 * k;
    const int y = p >> 3, x = p & 7;
    const float* it = patch + (y + 1) * 10 + 1 + x;
    refv[k] = it[0];
    if (!edgelet) {
      jx[k] = 0.5f * (it[1] - it[-1]);
      jy[k] = 0.5f * (it[10] - it[-10]);
      wgt[k] = sqrtf((float)(250.0 / (250.0 + (double)(jx[k] * jx[k] + jy[k] * jy[k]))));
      h00 += (jx[k] * jx[k]) * wgt[k]; h01 += (jx[k] * jy[k]) * wgt[k]; h02 += jx[k] * wgt[k];
      h11 += (jy[k] * jy[k]) * wgt[k]; h12 += jy[k] * wgt[k]; h22 += wgt[k];
    } else {
      jx[k] = (float)(0.5 * (double)(dirx * (it[1] - it[-1]) + diry * (it[10] - it[-10])));
      jy[k] = 0.f;
      wgt[k] = sqrtf((float)(250.0 / (250.0 + (double)(jx[k] * jx[k]))));
      h00 += (jx[k] * jx[k]) * wgt[k]; h02 += jx[k] * wgt[k]; h22 += wgt[k];
    }
  }
  h00 = warp_sum(h00); h02 = warp_sum(h02); h22 = warp_sum(h22);
  float Hinv[9];
  double h_inv_out = 0;
  if (!edgelet) {
    h01 = warp_sum(h01); h11 = warp_sum(h11); h12 = warp_sum(h12);
    const float k1 = (float)(1 + 0.001);
    float H[9] = {h00 * k1, h01, h02, h01, h11 * k1, h12, h02, h12, h22 * k1};
    inv3(H, Hinv);
  } else {
    const float k1 = (float)(1 + 0.001);
    const float H0 = h00 * k1, H1 = h02, H3 = h22 * k1;
    h_inv_out = 1.0 / (double)H0 * 8 * 8;  // h_inv = 1.0/H(0,0)*patch_size*patch_size (:207)
    const float det = H0 * H3 - H1 * H1;
    const float id = 1.0f / det;
    Hinv[0] = H3 * id; Hinv[1] = -H1 * id; Hinv[2] = -H1 * id; Hinv[3] = H0 * id;
  }

  // ---- iterations (feature_alignment.cpp:531-597 / :226-300) ------------------------------------------------------------------
  const uint8_t* cur = P.cur_pyr + P.g.off[sl];
  const int cols = P.g.w[sl], rows = P.g.h[sl];
  const double px_s0 = jb.px_cur[0] / (double)(1 << sl), px_s1 = jb.px_cur[1] / (double)(1 << sl);
  float u = (float)px_s0, v = (float)px_s1;
  const float min_update_squared = edgelet ? (float)(0.01 * 0.01) : (float)(0.03 * 0.03);
  float mean_diff = 0.f, chi2 = 0.f;
  float curv[2] = {0.f, 0.f};
  bool converged = false, nan_fail = false;
  for (int iter = 0; iter < P.max_iter; ++iter) {
    const int u_r = __float2int_rd(u), v_r = __float2int_rd(v);  // NaN -> 0 -> leaves through the bounds test like x86's INT_MIN
    if (u_r < 4 || v_r < 4 || u_r >= cols - 4 || v_r >= rows - 4) break;
    if (isnan(u) || isnan(v)) { nan_fail = true; break; }
    const float sx = u - (float)u_r, sy = v - (float)v_r;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy));
    const float wTR = (float)(sx * (1.0 - sy));
    const float wBL = (float)((1.0 - sx) * sy);
    const float wBR = sx * sy;
    float j0 = 0.f, j1 = 0.f, j2 = 0.f, nchi = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = lane + 32 * k;
      const int y = p >> 3, x = p & 7;
      const uint8_t* it = cur + (v_r + y - 4) * cols + u_r - 4 + x;
      const float sp = wTL * __ldg(it) + wTR * __ldg(it + 1) + wBL * __ldg(it + cols) + wBR * __ldg(it + cols + 1);
      const float res = sp - refv[k] + mean_diff;
      j0 -= res * jx[k] * wgt[k];
      if (!edgelet) j1 -= res * jy[k] * wgt[k];
      j2 -= res * wgt[k];
      nchi += res * res * wgt[k];
      curv[k] = sp;
    }
    j0 = warp_sum(j0); j2 = warp_sum(j2); nchi = warp_sum(nchi);
    chi2 = nchi;
    if (!edgelet) {
      j1 = warp_sum(j1);
      const float up0 = Hinv[0] * j0 + Hinv[1] * j1 + Hinv[2] * j2;
      const float up1 = Hinv[3] * j0 + Hinv[4] * j1 + Hinv[5] * j2;
      const float up2 = Hinv[6] * j0 + Hinv[7] * j1 + Hinv[8] * j2;
      u += up0; v += up1; mean_diff += up2;
      if (up0 * up0 + up1 * up1 < min_update_squared) { converged = true; break; }
    } else {
      const float up0 = Hinv[0] * j0 + Hinv[1] * j2;
      const float up1 = Hinv[2] * j0 + Hinv[3] * j2;
      u += up0 * dirx; v += up0 * diry; mean_diff += up1;
      if (up0 * up0 < min_update_squared) { converged = true; break; }
    }
  }
  if (chi2 > 1000.f * 64.f) converged = false;
  if (nan_fail) converged = false;
  // `return false` on NaN leaves cur_px_estimate untouched (:537); otherwise cur_px_estimate << u, v
  const double ps0 = nan_fail ? px_s0 : (double)u, ps1 = nan_fail ? px_s1 : (double)v;

  bool ok = converged;
  // ---- checkNormal (src/matcher.cpp:406-440), edgelets only ----------------------------------------------------------------
  if (ok && edgelet) {
    const float uf = (float)ps0, vf = (float)ps1;
    const int ui = __float2int_rd(uf), vi = __float2int_rd(vf);
    const float sx = uf - (float)ui, sy = vf - (float)vi;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy));
    const float wTR = (float)(sx * (1.0 - sy));
    const float wBL = (float)((1.0 - sx) * sy);
    const float wBR = (float)(1.0 - wTL - wTR - wBL);
    const int16_t* gxp = P.cur_sobel + P.sobel_off[sl];
    const int16_t* gyp = gxp + (size_t)cols * rows;
    const size_t o = (size_t)vi * cols + ui;
    const double nx = (double)wTL * gxp[o] + (double)wTR * gxp[o + 1] + (double)wBL * gxp[o + cols] + (double)wBR * gxp[o + cols + 1];
    const double ny = (double)wTL * gyp[o] + (double)wTR * gyp[o + 1] + (double)wBL * gyp[o + cols] + (double)wBR * gyp[o + cols + 1];
    const double nn = sqrt(nx * nx + ny * ny);
    ok = (dir_d0 * (nx / nn) + dir_d1 * (ny / nn)) > (double)0.86f;  // Config::edgeLetCosAngle (src/config.cpp:58)
  }
  // ---- checkNCC (src/matcher.cpp:379-404) -------------------------------------------------------------------------------------
  if (ok) {
    float mean1 = warp_sum(refv[0] + refv[1]), mean2 = warp_sum(curv[0] + curv[1]);
    mean1 /= 64.f; mean2 /= 64.f;
    float num = 0, d1 = 0, d2 = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float p1 = refv[k] - mean1, p2 = curv[k] - mean2;
      num += p1 * p2; d1 += p1 * p1; d2 += p2 * p2;
    }
    num = warp_sum(num); d1 = warp_sum(d1); d2 = warp_sum(d2);
    const float thresh = jb.ncc_thresh > 0.f ? jb.ncc_thresh : 0.7f;  // findMatchDirect 0.7 (matcher.cpp:366), findMatchSeed 0.8 (:510)
    ok = ((double)num / ((double)sqrtf(d1 * d2) + 1e-12)) > (double)thresh;
  }
  if (ok) {
    const double dx = px_s0 - ps0, dy = px_s1 - ps1;
    ok = sqrt(dx * dx + dy * dy) < 20;
  }
  if (lane == 0) {
    hso_align_result r;
    r.ok = ok ? 1 : 0;
    r.align_converged = converged ? 1 : 0;
    r.px_cur[0] = ps0 * (double)(1 << sl);
    r.px_cur[1] = ps1 * (double)(1 << sl);
    r.h_inv = h_inv_out;
    out[m] = r;
  }
}

cudaError_t launch_align(const PyrGeom& g, const uint8_t* cur_pyr, const int16_t* cur_sobel, const AlignJobDev* jobs_dev, int M, int max_iter,
                         hso_align_result* out_dev, cudaStream_t stream, uint64_t* launches) {
  AlignKParams P;
  P.g = g;
  P.cur_pyr = cur_pyr;
  P.cur_sobel = cur_sobel;
  size_t so = 0;
  for (int l = 0; l < 3; ++l) {
    P.sobel_off[l] = so;
    if (l < g.n_levels) so += (size_t)2 * g.w[l] * g.h[l];
  }
  P.max_iter = max_iter;
  P.M = M;
  k_align<<<(M + ALIGN_WARPS - 1) / ALIGN_WARPS, ALIGN_WARPS * 32, 0, stream>>>(P, jobs_dev, out_dev);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace hso
