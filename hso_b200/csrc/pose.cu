// Per-frame pose refinement on sm_100a: pose_optimizer::optimizeLevenbergMarquardt3rd (src/pose_optimizer.cpp:399-771) with
// MADScaleEstimator (src/vikit/robust_cost.cpp:65-74), HuberWeightFunction k=1.345 evaluated in float (:129-148),
// hso::getMedian (include/hso/vikit/math_utils.h:119-126) and Frame::jacobian_xyz2uv (include/hso/frame.h:192-212).
//
// One CTA per frame, the whole Levenberg-Marquardt loop on the device. fp64 like the reference (floats exactly where it uses
// floats). Per trial: every thread accumulates the 21 + 6 normal-equation entries of its features in registers, shuffle tree in
// the warp, fixed-order sum across warps (run-to-run deterministic), thread 0 damps/solves (pivoted LDL^T) and applies the SE3
// update; a second pass reduces the robust chi^2 of the trial pose. The MAD scales and the chi^2 medians are exact order
// statistics (radix select on the IEEE bit patterns), so they equal nth_element's result.
#include "hso_internal.h"

namespace hso {

constexpr int POSE_THREADS = 256;
constexpr int POSE_WARPS = POSE_THREADS / 32;
constexpr int PNRED = 28;  // 21 A + 6 b + 1 chi2

struct PoseShared {
  double warp_part[POSE_WARPS][PNRED];
  double tot[PNRED];
  double Rt[12];          // T_f_w used by the pass in flight
  Se3d T_cur, T_new;
  double A[36], b[6], dT[6];
  uint32_t hist[2048];
  unsigned long long sel_prefix;
  uint32_t sel_k, sel_n;
  int flag;
};

HSO_DEV float huber_value(float t) {  // robust_cost.cpp:141-148
  const float k = 1.345f;
  const float t_abs = fabsf(t);
  return t_abs < k ? 1.0f : k / t_abs;
}

// Exact k-th smallest (k = n/2) of the 64-bit keys of class `cls` (cls < 0: all). nbits = 32 or 64 significant bits.
HSO_DEV unsigned long long block_select(const unsigned long long* keys, const int8_t* cls_arr, int cls, int F, int nbits, PoseShared* s, uint32_t* n_out) {
  unsigned long long prefix = 0, mask = 0;
  int hi = nbits;
  bool first = true;
  while (hi > 0) {
    const int width = hi >= 11 ? 11 : hi;
    const int shift = hi - width;
    const unsigned long long dmask = (1ull << width) - 1;
    for (int j = threadIdx.x; j < 2048; j += blockDim.x) s->hist[j] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < F; i += blockDim.x) {
      if (cls >= 0 && cls_arr[i] != cls) continue;
      const unsigned long long key = keys[i];
      if ((key & mask) == prefix) atomicAdd(&s->hist[(uint32_t)((key >> shift) & dmask)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      uint32_t local = 0;
      for (int j = 0; j < 64; ++j) local += s->hist[lane * 64 + j];
      uint32_t incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      uint32_t k;
      if (first) { k = total / 2; if (lane == 0) s->sel_n = total; }
      else k = s->sel_k;
      __syncwarp();  // every lane has read sel_k before the lane that owns the bucket overwrites it (racecheck: read/write hazard inside the warp)
      const uint32_t excl = incl - local;
      if (total > 0 && k >= excl && k < incl) {
        uint32_t cum = excl;
        int d = lane * 64;
        for (int j = 0; j < 64; ++j) {
          const uint32_t c = s->hist[lane * 64 + j];
          if (k < cum + c) { d = lane * 64 + j; break; }
          cum += c;
        }
        s->sel_k = k - cum;
        s->sel_prefix = prefix | ((unsigned long long)d << shift);
      }
      if (total == 0 && lane == 0) { s->sel_k = 0; s->sel_prefix = 0; }
    }
    __syncthreads();
    prefix = s->sel_prefix;
    mask |= dmask << shift;
    hi = shift;
    first = false;
  }
  if (n_out) *n_out = s->sel_n;
  __syncthreads();
  return prefix;
}

HSO_DEV void block_reduce(const double* v, int n, PoseShared* s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < n; ++k) {
    const double x = warp_sum(v[k]);
    if (lane == 0) s->warp_part[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < n) {
    double sum = 0;
    for (int w = 0; w < POSE_WARPS; ++w) sum += s->warp_part[w][threadIdx.x];
    s->tot[threadIdx.x] = sum;
  }
  __syncthreads();
}

struct Feat {
  double e0, e1;    // scaled reprojection error
  double px, py, pz;  // pTarget
};

// pTarget = (T_f_w * T_host^-1) * pHost ; e = (project2d(f) - project2d(pTarget)) / 2^level  (pose_optimizer.cpp:431-440)
HSO_DEV Feat feat_residual(const PoseJobDev& j, const double* Tth /*[K][12]*/, int i) {
  const double* T = Tth + 12 * j.host_idx[i];
  const double hx = j.p_host[3 * i], hy = j.p_host[3 * i + 1], hz = j.p_host[3 * i + 2];
  Feat r;
  r.px = T[0] * hx + T[1] * hy + T[2] * hz + T[3];
  r.py = T[4] * hx + T[5] * hy + T[6] * hz + T[7];
  r.pz = T[8] * hx + T[9] * hy + T[10] * hz + T[11];
  const double fx = j.f[3 * i], fy = j.f[3 * i + 1], fz = j.f[3 * i + 2];
  const double s = 1.0 / (double)(1 << j.level[i]);
  r.e0 = (fx / fz - r.px / r.pz) * s;
  r.e1 = (fy / fz - r.py / r.pz) * s;
  return r;
}

HSO_DEV void compute_Tth(const PoseJobDev& j, const Se3d& T, const Se3d* T_host_inv, double* Tth) {
  for (int k = threadIdx.x; k < j.K; k += blockDim.x) se3_to_rt(se3_mul(T, T_host_inv[k]), Tth + 12 * k);
  __syncthreads();
}

HSO_DEV double robust_chi2_local(const PoseJobDev& j, const double* Tth, float scale_pt, float scale_ls) {
  double c = 0;
  for (int i = threadIdx.x; i < j.F; i += blockDim.x) {
    const Feat r = feat_residual(j, Tth, i);
    if (j.ftype[i] == 1) {
      const double e = j.grad[2 * i] * r.e0 + j.grad[2 * i + 1] * r.e1;
      double w = (double)huber_value((float)(fabs(e) / (double)scale_ls));
      if (j.ptype[i] == 1) w *= 0.5;
      c += e * e * w;
    } else {
      const double e = sqrt(r.e0 * r.e0 + r.e1 * r.e1);
      double w = (double)huber_value((float)(e / (double)scale_pt));
      if (j.ptype[i] == 1) w *= 0.5;
      c += e * e * w;
    }
  }
  return c;
}

__device__ __noinline__ void inv6(const double* A, double* Ai) {  // Gauss-Jordan with partial pivoting
  double m[6][12];
  for (int i = 0; i < 6; ++i)
    for (int k = 0; k < 6; ++k) { m[i][k] = A[i * 6 + k]; m[i][6 + k] = (i == k) ? 1.0 : 0.0; }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
    if (p != c) for (int k = 0; k < 12; ++k) { const double t = m[c][k]; m[c][k] = m[p][k]; m[p][k] = t; }
    const double d = m[c][c];
    for (int k = 0; k < 12; ++k) m[c][k] /= d;
    for (int r = 0; r < 6; ++r) {
      if (r == c) continue;
      const double fct = m[r][c];
      if (fct != 0.0) for (int k = 0; k < 12; ++k) m[r][k] -= fct * m[c][k];
    }
  }
  for (int i = 0; i < 6; ++i) for (int k = 0; k < 6; ++k) Ai[i * 6 + k] = m[i][6 + k];
}

__global__ void __launch_bounds__(POSE_THREADS) k_pose_lm(const PoseJobDev* __restrict__ jobs, const PoseScratch* __restrict__ scr, double reproj_thresh,
                                                            int n_iter, double err_mult2) {
  __shared__ PoseShared s;
  const PoseJobDev j = jobs[blockIdx.x];
  const PoseScratch sc = scr[blockIdx.x];
  hso_pose_result* out = j.out;
  const int tid = threadIdx.x;

  if (tid == 0) {
    s.T_cur = se3_from_rt(j.T_f_w_in);
    out->n_trials_total = 0; out->early_return = 0;
    se3_to_rt(s.T_cur, out->T_f_w);
    for (int i = 0; i < 36; ++i) out->cov[i] = 0;
    out->estimated_scale = out->error_init = out->error_final = 0;
    out->num_obs = 0; out->error_in_px = 0;
  }
  for (int i = tid; i < j.F; i += blockDim.x) j.outlier[i] = 0;
  for (int k = tid; k < j.K; k += blockDim.x) sc.T_host_inv[k] = se3_inverse(se3_from_rt(j.T_host_w + 12 * k));
  __syncthreads();
  if (j.F == 0) {  // errors_pt.empty() && errors_ls.empty() -> return (pose_optimizer.cpp:456)
    if (tid == 0) out->early_return = 1;
    return;
  }

  // ---- pass 0: errors for the scale estimate (:426-454) ---------------------------------------------------------------------
  compute_Tth(j, s.T_cur, sc.T_host_inv, sc.Tth);
  for (int i = tid; i < j.F; i += blockDim.x) {
    const Feat r = feat_residual(j, sc.Tth, i);
    float err;
    if (j.ftype[i] == 1) {
      const float error_ls = (float)(j.grad[2 * i] * r.e0 + j.grad[2 * i + 1] * r.e1);
      err = fabsf(error_ls);
      sc.cls[i] = 1;
    } else {
      err = (float)sqrt(r.e0 * r.e0 + r.e1 * r.e1);
      sc.cls[i] = 0;
    }
    sc.keys[i] = (unsigned long long)__float_as_uint(err);
  }
  __syncthreads();
  uint32_t n_pt = 0, n_ls = 0, n_all = 0;
  const float med_pt = __uint_as_float((uint32_t)block_select(sc.keys, sc.cls, 0, j.F, 32, &s, &n_pt));
  const float med_ls = __uint_as_float((uint32_t)block_select(sc.keys, sc.cls, 1, j.F, 32, &s, &n_ls));
  const float med_all = __uint_as_float((uint32_t)block_select(sc.keys, sc.cls, -1, j.F, 32, &s, &n_all));
  float scale_pt = 0.f, scale_ls = 0.f;
  if (n_pt > 0 && n_ls > 0) { scale_pt = 1.4826f * med_pt; scale_ls = 1.4826f * med_ls; }
  else if (n_pt > 0) { scale_pt = 1.4826f * med_pt; scale_ls = (float)(0.5 * (double)scale_pt); }
  else { scale_ls = 1.4826f * med_ls; scale_pt = 2.f * scale_ls; }
  // chi2_vec_init holds float squares of these errors; squaring is monotone, so its median is the square of the median error
  const double error_init = sqrt((double)(med_all * med_all)) * err_mult2;

  // ---- initial robust chi2 (:488-526) ------------------------------------------------------------------------------------------
  double acc[PNRED];
  acc[0] = robust_chi2_local(j, sc.Tth, scale_pt, scale_ls);
  block_reduce(acc, 1, &s);
  double chi2 = s.tot[0];
  double mu = 0.1, nu = 2.0;
  bool stop = false;
  int n_trials_total = 0;
  __syncthreads();

  // ---- LM iterations (:531-689) --------------------------------------------------------------------------------------------------
  for (int iter = 0; iter < n_iter && !stop; ++iter) {
    double rho = 0;
    int n_trials = 0;
    do {
      for (int k = 0; k < PNRED; ++k) acc[k] = 0;
      for (int i = tid; i < j.F; i += blockDim.x) {
        const Feat r = feat_residual(j, sc.Tth, i);
        const double zi = 1.0 / r.pz, zi2 = zi * zi;
        const double sic = 1.0 / (double)(1 << j.level[i]);
        double J0[6], J1[6];
        J0[0] = -zi; J0[1] = 0.0; J0[2] = r.px * zi2; J0[3] = r.py * J0[2]; J0[4] = -(1.0 + r.px * J0[2]); J0[5] = r.py * zi;
        J1[0] = 0.0; J1[1] = -zi; J1[2] = r.py * zi2; J1[3] = 1.0 + r.py * J1[2]; J1[4] = -J0[3]; J1[5] = -r.px * zi;
#pragma unroll
        for (int k = 0; k < 6; ++k) { J0[k] *= sic; J1[k] *= sic; }
        if (j.ftype[i] == 1) {
          const double gx = j.grad[2 * i], gy = j.grad[2 * i + 1];
          double Je[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) Je[k] = gx * J0[k] + gy * J1[k];
          const double e_edge = gx * r.e0 + gy * r.e1;
          double w = (double)huber_value((float)(fabs(e_edge) / (double)scale_ls));
          if (j.ptype[i] == 1) w *= 0.5;
          int idx = 0;
#pragma unroll
          for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int c = a; c < 6; ++c) acc[idx++] += Je[a] * Je[c] * w;
            acc[21 + a] -= Je[a] * e_edge * w;
          }
        } else {
          double w = (double)huber_value((float)(sqrt(r.e0 * r.e0 + r.e1 * r.e1) / (double)scale_pt));
          if (j.ptype[i] == 1) w *= 0.5;
          int idx = 0;
#pragma unroll
          for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int c = a; c < 6; ++c) acc[idx++] += (J0[a] * J0[c] + J1[a] * J1[c]) * w;
            acc[21 + a] -= (J0[a] * r.e0 + J1[a] * r.e1) * w;
          }
        }
      }
      block_reduce(acc, 27, &s);
      if (tid == 0) {
        int idx = 0;
        for (int a = 0; a < 6; ++a)
          for (int c = a; c < 6; ++c) { s.A[a * 6 + c] = s.A[c * 6 + a] = s.tot[idx]; ++idx; }
        for (int a = 0; a < 6; ++a) s.b[a] = s.tot[21 + a];
        for (int a = 0; a < 6; ++a) s.A[a * 6 + a] += s.A[a * 6 + a] * mu;  // A += (A.diagonal()*mu).asDiagonal() (:594)
        ldlt_solve<6>(s.A, s.b, s.dT);
        s.flag = isnan(s.dT[0]) ? 0 : 1;
        if (s.flag) s.T_new = se3_mul(se3_exp(s.dT), s.T_cur);
      }
      __syncthreads();
      ++n_trials_total;
      double new_chi2 = 0.0;
      if (s.flag) {
        compute_Tth(j, s.T_new, sc.T_host_inv, sc.Tth);
        acc[0] = robust_chi2_local(j, sc.Tth, scale_pt, scale_ls);
        block_reduce(acc, 1, &s);
        new_chi2 = s.tot[0];
        rho = chi2 - new_chi2;
      } else {
        rho = -1;
      }
      if (rho > 0) {
        if (tid == 0) s.T_cur = s.T_new;
        chi2 = new_chi2;
        double nm = -1;
        for (int k = 0; k < 6; ++k) nm = fmax(nm, fabs(s.dT[k]));
        stop = nm <= 0.0000000001;  // EPS, include/hso/global.h:105
        mu *= fmax(1. / 3., fmin(1. - pow(2 * rho - 1, 3.0), 2. / 3.));
        nu = 2.;
      } else {
        mu *= nu;
        nu *= 2.;
        if (mu < 0.0001) mu = 0.0001;
        ++n_trials;
        if (n_trials >= 5) stop = true;
        __syncthreads();
        if (s.flag) compute_Tth(j, s.T_cur, sc.T_host_inv, sc.Tth);  // back to the accepted pose
      }
      __syncthreads();
    } while (!(rho > 0 || stop));
  }

  // ---- covariance, outlier culling, medians (:691-767) -----------------------------------------------------------------------------
  if (tid == 0) {
    double As[36];
    const double s2 = err_mult2 * err_mult2;
    for (int i = 0; i < 36; ++i) As[i] = (n_trials_total > 0 ? s.A[i] : 0.0) * s2;
    inv6(As, out->cov);
  }
  const float thr_pt = (j.n_fts_total < 80) ? (float)(sqrt(5.991) / err_mult2) : (float)(reproj_thresh / err_mult2);
  const float thr_ls = (float)(1.3 / err_mult2);
  int n_deleted = 0;
  for (int i = tid; i < j.F; i += blockDim.x) {
    const Feat r = feat_residual(j, sc.Tth, i);
    if (j.ftype[i] == 1) {
      const double error_ls = j.grad[2 * i] * r.e0 + j.grad[2 * i + 1] * r.e1;
      if (fabs(error_ls) > (double)thr_ls) { ++n_deleted; j.outlier[i] = 1; }
      sc.keys[i] = (unsigned long long)__double_as_longlong(error_ls * error_ls);
    } else {
      const float error_pt = (float)sqrt(r.e0 * r.e0 + r.e1 * r.e1);
      if (error_pt > thr_pt) { ++n_deleted; j.outlier[i] = 1; }
      sc.keys[i] = (unsigned long long)__double_as_longlong((double)(error_pt * error_pt));
    }
  }
  acc[0] = (double)n_deleted;
  block_reduce(acc, 1, &s);
  const double deleted = s.tot[0];
  __syncthreads();
  uint32_t n_fin = 0;
  const double med_final = __longlong_as_double((long long)block_select(sc.keys, sc.cls, -1, j.F, 64, &s, &n_fin));
  if (tid == 0) {
    const double error_final = sqrt(med_final) * err_mult2;
    se3_to_rt(s.T_cur, out->T_f_w);
    out->estimated_scale = (double)scale_pt * err_mult2;
    out->error_init = error_init;
    out->error_final = error_final;
    out->num_obs = (uint64_t)((long long)(n_pt + n_ls) - (long long)deleted);
    out->error_in_px = error_final < 1.5 ? 1.0f : (float)(1.5 / error_final);
    out->n_trials_total = n_trials_total;
  }
}

cudaError_t launch_pose(const PoseJobDev* jobs_dev, const PoseScratch* scratch_dev, int B, double reproj_thresh, int n_iter, double err_mult2,
                        cudaStream_t stream, uint64_t* launches) {
  k_pose_lm<<<B, POSE_THREADS, 0, stream>>>(jobs_dev, scratch_dev, reproj_thresh, n_iter, err_mult2);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace hso
