// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of pose_optimizer::optimizeLevenbergMarquardt3rd
// (src/pose_optimizer.cpp:399-771), the only live pose optimiser (call site src/frame_handler_mono.cpp:241-243),
// with robust_cost::MADScaleEstimator (src/vikit/robust_cost.cpp:65-74), HuberWeightFunction k=1.345 in float
// (:129-148), hso::getMedian / norm_max / project2d (include/hso/vikit/math_utils.h:88-126) and
// Frame::jacobian_xyz2uv (include/hso/frame.h:192-212). Double throughout except where the reference uses float.
// Quirks kept: rho is the plain chi2 difference; Cov_ uses the last *damped* A; the <80-feature threshold counts all
// features (with or without point); culled features still contribute to the final median.
// Parity status: unpinned by the reference (no tests/golden vectors).
#include <cmath>
#include <vector>

#include "hso_oracle.h"
#include "oracle_math.hpp"

using namespace orc;

namespace {

const double EPS = 0.0000000001;  // include/hso/global.h:105

inline float huber_value(float t) {  // robust_cost.cpp:141-148, k = 1.345f
  const float k = 1.345f;
  const float t_abs = std::abs(t);
  if (t_abs < k) return 1.0f;
  return k / t_abs;
}
inline float mad_scale(std::vector<float>& errors) {  // robust_cost.cpp:67-74
  return 1.4826f * median_inplace(errors);
}
inline void jacobian_xyz2uv(const V3& p, double J[2][6]) {
  const double x = p.x, y = p.y;
  const double z_inv = 1. / p.z;
  const double z_inv_2 = z_inv * z_inv;
  J[0][0] = -z_inv; J[0][1] = 0.0; J[0][2] = x * z_inv_2; J[0][3] = y * J[0][2]; J[0][4] = -(1.0 + x * J[0][2]); J[0][5] = y * z_inv;
  J[1][0] = 0.0; J[1][1] = -z_inv; J[1][2] = y * z_inv_2; J[1][3] = 1.0 + y * J[1][2]; J[1][4] = -J[0][3]; J[1][5] = -x * z_inv;
}

// General 6x6 inverse for Cov_ (Eigen uses partial-pivot LU for dynamic/large fixed sizes).
void inv6(const double* A, double* Ai) {
  double m[6][12];
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) { m[i][j] = A[i * 6 + j]; m[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r) if (std::fabs(m[r][c]) > std::fabs(m[p][c])) p = r;
    if (p != c) for (int j = 0; j < 12; ++j) std::swap(m[c][j], m[p][j]);
    double d = m[c][c];
    for (int j = 0; j < 12; ++j) m[c][j] /= d;
    for (int r = 0; r < 6; ++r) {
      if (r == c) continue;
      double fct = m[r][c];
      if (fct != 0.0) for (int j = 0; j < 12; ++j) m[r][j] -= fct * m[c][j];
    }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Ai[i * 6 + j] = m[i][6 + j];
}

struct Obs {
  V3 f, pHost;
  int host;
  double gx, gy;
  int level, ftype, ptype;
};

inline void residual(const Obs& o, const SE3& Tth, V3* pTarget, double e[2]) {
  *pTarget = Tth.apply(o.pHost);
  e[0] = o.f.x / o.f.z - pTarget->x / pTarget->z;  // project2d(f) - project2d(pTarget)
  e[1] = o.f.y / o.f.z - pTarget->y / pTarget->z;
  const double s = 1.0 / (1 << o.level);
  e[0] *= s;
  e[1] *= s;
}

}  // namespace

extern "C" void orc_pose_optimize(double reproj_thresh, int n_iter, double err_mult2, int n_fts_total, int F, const double* f,
                                  const double* p_host, const int32_t* host_idx, const double* T_host_w, const double* grad,
                                  const int8_t* level, const int8_t* ftype, const int8_t* ptype, const double T_f_w_in[12],
                                  uint8_t* outlier_out, orc_pose_result* out) {
  double chi2 = 0.0, rho = 0, mu = 0.1, nu = 2.0;
  bool stop = false;
  int n_trials = 0;
  const int n_trials_max = 5;
  out->n_trials_total = 0;
  out->early_return = 0;
  SE3 T_f_w = SE3::from_rt(T_f_w_in);
  T_f_w.to_rt(out->T_f_w);
  for (int i = 0; i < 36; ++i) out->cov[i] = 0;
  out->estimated_scale = out->error_init = out->error_final = 0;
  out->num_obs = 0;
  out->error_in_px = 0;
  for (int i = 0; i < F; ++i) outlier_out[i] = 0;

  int K = 0;
  for (int i = 0; i < F; ++i) K = std::max(K, host_idx[i] + 1);
  std::vector<SE3> T_host_inv(K);
  for (int k = 0; k < K; ++k) T_host_inv[k] = SE3::from_rt(T_host_w + 12 * k).inverse();
  std::vector<Obs> obs(F);
  for (int i = 0; i < F; ++i) {
    obs[i].f = {f[3 * i], f[3 * i + 1], f[3 * i + 2]};
    obs[i].pHost = {p_host[3 * i], p_host[3 * i + 1], p_host[3 * i + 2]};
    obs[i].host = host_idx[i];
    obs[i].gx = grad[2 * i]; obs[i].gy = grad[2 * i + 1];
    obs[i].level = level[i]; obs[i].ftype = ftype[i]; obs[i].ptype = ptype[i];
  }

  std::vector<double> chi2_vec_init, chi2_vec_final;
  std::vector<float> errors_pt, errors_ls;
  double A[36], b[6];
  for (int i = 0; i < 36; ++i) A[i] = 0;
  for (int i = 0; i < 6; ++i) b[i] = 0;

  // pass 0: residuals for the scale estimate (pose_optimizer.cpp:426-454)
  for (const Obs& o : obs) {
    SE3 Tth = T_f_w.mul(T_host_inv[o.host]);
    V3 pT; double e[2];
    residual(o, Tth, &pT, e);
    if (o.ftype == 1) {
      float error_ls = o.gx * e[0] + o.gy * e[1];
      errors_ls.push_back(std::fabs(error_ls));
      chi2_vec_init.push_back(error_ls * error_ls);
    } else {
      float error_pt = std::sqrt(e[0] * e[0] + e[1] * e[1]);
      errors_pt.push_back(error_pt);
      chi2_vec_init.push_back(error_pt * error_pt);
    }
  }
  if (errors_pt.empty() && errors_ls.empty()) { out->early_return = 1; return; }

  float estimated_scale_pt = 0, estimated_scale_ls = 0;
  if (!errors_pt.empty() && !errors_ls.empty()) {
    estimated_scale_pt = mad_scale(errors_pt);
    estimated_scale_ls = mad_scale(errors_ls);
  } else if (!errors_pt.empty() && errors_ls.empty()) {
    estimated_scale_pt = mad_scale(errors_pt);
    estimated_scale_ls = 0.5 * estimated_scale_pt;
  } else if (errors_pt.empty() && !errors_ls.empty()) {
    estimated_scale_ls = mad_scale(errors_ls);
    estimated_scale_pt = 2 * estimated_scale_ls;
  }
  double estimated_scale = estimated_scale_pt;

  auto robust_chi2 = [&](const SE3& T) {
    double c = 0;
    for (const Obs& o : obs) {
      SE3 Tth = T.mul(T_host_inv[o.host]);
      V3 pT; double e[2];
      residual(o, Tth, &pT, e);
      if (o.ftype == 1) {
        double error_ls = o.gx * e[0] + o.gy * e[1];
        double weight = huber_value(std::fabs(error_ls) / estimated_scale_ls);
        if (o.ptype == 1) weight *= 0.5;
        c += error_ls * error_ls * weight;
      } else {
        double error_pt = std::sqrt(e[0] * e[0] + e[1] * e[1]);
        double weight = huber_value(error_pt / estimated_scale_pt);
        if (o.ptype == 1) weight *= 0.5;
        c += error_pt * error_pt * weight;
      }
    }
    return c;
  };

  chi2 = robust_chi2(T_f_w);  // pose_optimizer.cpp:488-526
  uint64_t num_obs = errors_pt.size() + errors_ls.size();

  for (int iter = 0; iter < n_iter; iter++) {
    rho = 0;
    n_trials = 0;
    do {
      SE3 T_new;
      double new_chi2 = 0.0;
      for (int i = 0; i < 36; ++i) A[i] = 0;
      for (int i = 0; i < 6; ++i) b[i] = 0;
      for (const Obs& o : obs) {
        SE3 Tth = T_f_w.mul(T_host_inv[o.host]);
        V3 pT; double e[2];
        residual(o, Tth, &pT, e);
        double J[2][6];
        jacobian_xyz2uv(pT, J);
        double sqrt_inv_cov = 1.0 / (1 << o.level);
        for (int k = 0; k < 6; ++k) { J[0][k] *= sqrt_inv_cov; J[1][k] *= sqrt_inv_cov; }
        if (o.ftype == 1) {
          double Je[6];
          for (int k = 0; k < 6; ++k) Je[k] = o.gx * J[0][k] + o.gy * J[1][k];
          double e_edge = o.gx * e[0] + o.gy * e[1];
          double weight = huber_value(std::fabs(e_edge) / estimated_scale_ls);
          if (o.ptype == 1) weight *= 0.5;
          for (int r = 0; r < 6; ++r) {
            for (int c = 0; c < 6; ++c) A[r * 6 + c] += Je[r] * Je[c] * weight;
            b[r] -= Je[r] * e_edge * weight;
          }
        } else {
          double weight = huber_value(std::sqrt(e[0] * e[0] + e[1] * e[1]) / estimated_scale_pt);
          if (o.ptype == 1) weight *= 0.5;
          for (int r = 0; r < 6; ++r) {
            for (int c = 0; c < 6; ++c) A[r * 6 + c] += (J[0][r] * J[0][c] + J[1][r] * J[1][c]) * weight;
            b[r] -= (J[0][r] * e[0] + J[1][r] * e[1]) * weight;
          }
        }
      }
      for (int i = 0; i < 6; ++i) A[i * 6 + i] += A[i * 6 + i] * mu;  // A += (A.diagonal()*mu).asDiagonal()
      double dT[6];
      ldlt_solve<6>(A, b, dT);
      out->n_trials_total++;
      if (!std::isnan(dT[0])) {
        T_new = SE3::exp(dT).mul(T_f_w);
        new_chi2 = robust_chi2(T_new);
        rho = chi2 - new_chi2;
      } else {
        rho = -1;
      }
      if (rho > 0) {
        T_f_w = T_new;
        chi2 = new_chi2;
        double nm = -1;
        for (int k = 0; k < 6; ++k) nm = std::max(nm, std::fabs(dT[k]));
        stop = nm <= EPS;
        mu *= std::max(1. / 3., std::min(1. - std::pow(2 * rho - 1, 3), 2. / 3.));
        nu = 2.;
      } else {
        mu *= nu;
        nu *= 2.;
        if (mu < 0.0001) mu = 0.0001;
        ++n_trials;
        if (n_trials >= n_trials_max) stop = true;
      }
    } while (!(rho > 0 || stop));
    if (stop) break;
  }

  // Cov_ = pixel_variance * (A * errMult2^2)^-1 with the last trial's damped A (pose_optimizer.cpp:691-692)
  {
    double As[36];
    const double s2 = std::pow(err_mult2, 2);
    for (int i = 0; i < 36; ++i) As[i] = A[i] * s2;
    inv6(As, out->cov);
  }
  const float reproj_thresh_scaled_pt = (n_fts_total < 80) ? sqrt(5.991) / err_mult2 : reproj_thresh / err_mult2;
  const float reproj_thresh_scaled_ls = 1.3 / err_mult2;
  size_t n_deleted_refs = 0;
  for (int i = 0; i < F; ++i) {
    const Obs& o = obs[i];
    SE3 Tth = T_f_w.mul(T_host_inv[o.host]);
    V3 pT; double e[2];
    residual(o, Tth, &pT, e);
    if (o.ftype == 1) {
      double error_ls = o.gx * e[0] + o.gy * e[1];
      if (std::fabs(error_ls) > reproj_thresh_scaled_ls) { ++n_deleted_refs; outlier_out[i] = 1; }
      chi2_vec_final.push_back(error_ls * error_ls);
    } else {
      float error_pt = std::sqrt(e[0] * e[0] + e[1] * e[1]);
      if (error_pt > reproj_thresh_scaled_pt) { ++n_deleted_refs; outlier_out[i] = 1; }
      chi2_vec_final.push_back(error_pt * error_pt);
    }
  }
  double error_init = 0.0, error_final = 0.0;
  if (!chi2_vec_init.empty()) error_init = std::sqrt(median_inplace(chi2_vec_init)) * err_mult2;
  if (!chi2_vec_final.empty()) error_final = std::sqrt(median_inplace(chi2_vec_final)) * err_mult2;
  estimated_scale *= err_mult2;
  num_obs -= n_deleted_refs;

  T_f_w.to_rt(out->T_f_w);
  out->estimated_scale = estimated_scale;
  out->error_init = error_init;
  out->error_final = error_final;
  out->num_obs = num_obs;
  out->error_in_px = error_final < 1.5 ? 1.0 : 1.5 / error_final;
}
