// ORACLE — TEST INFRASTRUCTURE ONLY. Line-faithful, single-threaded CPU restatement of hso::CoarseTracker
// (src/CoarseTracker.cpp:51-646, include/hso/CoarseTracker.h:58-122) and Accumulator7
// (include/hso/MatrixAccumulator.h:29-141). Same float/double mix, same loop order, same quirks:
//   * pattern idx2 lists {-1,0} twice and lacks {0,-1}                      (CoarseTracker.h:69)
//   * reference weights use w_br = 1-(tl+tr+bl), current weights w_br = su*sv (CoarseTracker.cpp:467 vs :323)
//   * top level never saturates and uses E += hw r^2 ; lower levels hw r^2 (2-hw) (:350-361)
//   * H accumulated in float through the 3-tier accumulator, b in double     (:499-525)
//   * Jacobians ignore lens distortion while the projection applies it       (:252-253,302,372)
// Parity status: the reference ships no test/golden vector for this path and cannot be compiled
// here (needs Eigen/OpenCV/Boost) => "parity unpinned" by the reference; see DESIGN.md.
#include <cmath>
#include <cstdio>
#include <vector>

#include "hso_oracle.h"
#include "oracle_math.hpp"

using namespace orc;

namespace {

// include/hso/CoarseTracker.h:58-120 — only the first staticPatternNum entries of each row are ever read.
const int kPatternNum[8] = {1, 5, 9, 13, 13, 21, 25, 25};
const int kPatternPadding[8] = {1, 1, 1, 2, 2, 3, 2, 4};
const int kPattern[8][25][2] = {
    {{0, 0}},
    {{0, -1}, {-1, 0}, {0, 0}, {1, 0}, {0, 1}},
    {{-1, -1}, {-1, 0}, {-1, 1}, {-1, 0}, {0, 0}, {0, 1}, {1, -1}, {1, 0}, {1, 1}},
    {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {1, 1}, {0, 2}, {0, -1}, {-1, 0}, {1, 0}, {0, 1}},
    {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {1, 1}, {0, 2}, {-2, -2}, {-2, 2}, {2, -2}, {2, 2}},
    {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {1, 1}, {0, 2}, {-2, -2}, {-2, 2},
     {2, -2}, {2, 2}, {-3, -1}, {-3, 1}, {3, -1}, {3, 1}, {1, -3}, {-1, -3}, {1, 3}, {-1, 3}},
    {{-2, -2}, {-2, -1}, {-2, 0}, {-2, 1}, {-2, 2}, {-1, -2}, {-1, -1}, {-1, 0}, {-1, 1}, {-1, 2}, {0, -2}, {0, -1}, {0, 0},
     {0, 1}, {0, 2}, {1, -2}, {1, -1}, {1, 0}, {1, 1}, {1, 2}, {2, -2}, {2, -1}, {2, 0}, {2, 1}, {2, 2}},
    {{-4, -4}, {-4, -2}, {-4, 0}, {-4, 2}, {-4, 4}, {-2, -4}, {-2, -2}, {-2, 0}, {-2, 2}, {-2, 4}, {0, -4}, {0, -2}, {0, 0},
     {0, 2}, {0, 4}, {2, -4}, {2, -2}, {2, 0}, {2, 2}, {2, 4}, {4, -4}, {4, -2}, {4, 0}, {4, 2}, {4, 4}},
};
const int kPatternOffset = 2;  // CoarseTracker.h:122

// DSO-style tiered float accumulator, include/hso/MatrixAccumulator.h:29-141. Only SSE lane 0 is ever written by
// updateSingleWeighted(off=0); the other three lanes stay 0 and the final 4-lane sum adds exact zeros.
struct Acc7 {
  float d[28], d1k[28], d1m[28];
  float numIn1, numIn1k, numIn1m;
  void initialize() {
    std::memset(d, 0, sizeof d); std::memset(d1k, 0, sizeof d1k); std::memset(d1m, 0, sizeof d1m);
    numIn1 = numIn1k = numIn1m = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < 28; ++i) d1k[i] = d[i] + d1k[i];
      numIn1k += numIn1; numIn1 = 0;
      std::memset(d, 0, sizeof d);
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < 28; ++i) d1m[i] = d1k[i] + d1m[i];
      numIn1m += numIn1k; numIn1k = 0;
      std::memset(d1k, 0, sizeof d1k);
    }
  }
  void update(float J0, float J1, float J2, float J3, float J4, float J5, float J6, float w) {
    float* pt = d;
    *pt += J0 * J0 * w; pt++; J0 *= w;
    *pt += J1 * J0; pt++; *pt += J2 * J0; pt++; *pt += J3 * J0; pt++; *pt += J4 * J0; pt++; *pt += J5 * J0; pt++; *pt += J6 * J0; pt++;
    *pt += J1 * J1 * w; pt++; J1 *= w;
    *pt += J2 * J1; pt++; *pt += J3 * J1; pt++; *pt += J4 * J1; pt++; *pt += J5 * J1; pt++; *pt += J6 * J1; pt++;
    *pt += J2 * J2 * w; pt++; J2 *= w;
    *pt += J3 * J2; pt++; *pt += J4 * J2; pt++; *pt += J5 * J2; pt++; *pt += J6 * J2; pt++;
    *pt += J3 * J3 * w; pt++; J3 *= w;
    *pt += J4 * J3; pt++; *pt += J5 * J3; pt++; *pt += J6 * J3; pt++;
    *pt += J4 * J4 * w; pt++; J4 *= w;
    *pt += J5 * J4; pt++; *pt += J6 * J4; pt++;
    *pt += J5 * J5 * w; pt++; J5 *= w;
    *pt += J6 * J5; pt++;
    *pt += J6 * J6 * w; pt++;
    numIn1++;
    shiftUp(false);
  }
  void finish(double H[49]) {
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 7; r++)
      for (int c = r; c < 7; c++) {
        float v = d1m[idx] + 0.f + 0.f + 0.f;
        H[r * 7 + c] = H[c * 7 + r] = (double)v;
        idx++;
      }
  }
};

// Frame::jacobian_xyz2uv — include/hso/frame.h:192-212 (negated projection Jacobian on the unit plane).
inline void jacobian_xyz2uv(const V3& p, double J[2][6]) {
  const double x = p.x, y = p.y;
  const double z_inv = 1. / p.z;
  const double z_inv_2 = z_inv * z_inv;
  J[0][0] = -z_inv; J[0][1] = 0.0; J[0][2] = x * z_inv_2; J[0][3] = y * J[0][2]; J[0][4] = -(1.0 + x * J[0][2]); J[0][5] = y * z_inv;
  J[1][0] = 0.0; J[1][1] = -z_inv; J[1][2] = y * z_inv_2; J[1][3] = 1.0 + y * J[1][2]; J[1][4] = -J[0][3]; J[1][5] = -x * z_inv;
}

struct Term { double J[7]; double w; double r; };

struct Tracker {
  const orc_cam* cam;
  bool inverse_comp;
  int max_level, min_level, n_iter;
  int F;
  const double *px, *f, *dist;

  // per-level state
  int level = 0, offset_all = 0, HALF_PATCH_SIZE = 2, PATCH_AREA = 13;
  const uint8_t *ref_img = nullptr, *cur_img = nullptr;
  int cols = 0, rows = 0;
  std::vector<float> ref_patch_cache;   // F x PATCH_AREA
  std::vector<char> visible;
  std::vector<double> jac_raw;          // 6 x (F*PATCH_AREA), column-major like the reference
  std::vector<Term> buf;
  int total_terms = 0, saturated_terms = 0;
  float huber_thresh = 0, outlier_thresh = 0;
  Acc7 acc;

  void set_level(int lvl, const uint8_t* r, const uint8_t* c, int w, int h) {
    level = lvl; ref_img = r; cur_img = c; cols = w; rows = h;
    offset_all = max_level - level + kPatternOffset;      // CoarseTracker.cpp:80
    HALF_PATCH_SIZE = kPatternPadding[offset_all];
    PATCH_AREA = kPatternNum[offset_all];
    ref_patch_cache.assign((size_t)F * PATCH_AREA, 0.f);
    visible.assign(F, 0);
    jac_raw.assign((size_t)6 * F * PATCH_AREA, 0.0);
  }

  // CoarseTracker.cpp:416-497
  void precomputeReferencePatches() {
    const int border = HALF_PATCH_SIZE + 1;
    const int stride = cols;
    const float scale = 1.0f / (1 << level);
    const double fxl = cam->fx * scale;
    const double fyl = cam->fy * scale;
    for (int i = 0; i < F; ++i) {
      // Feature::point == NULL (CoarseTracker.cpp:433) is carried by dist < 0 in the flattened interface: makeDepthRef
      // leaves -1 for such features (:212,:217) and every later stage skips dist < 0 (:290,:455,:557), so whether
      // `visible` is set for a feature with a point but p_ref.z < 1e-5 is unobservable.
      if (dist[i] < 0) continue;
      float u_ref = px[2 * i] * scale;
      float v_ref = px[2 * i + 1] * scale;
      int u_ref_i = floorf(u_ref);
      int v_ref_i = floorf(v_ref);
      if (u_ref_i - border < 0 || v_ref_i - border < 0 || u_ref_i + border >= cols || v_ref_i + border >= rows) continue;
      visible[i] = 1;
      double frame_jac[2][6] = {};
      if (inverse_comp) {
        double d = dist[i];
        if (d < 0) continue;
        V3 xyz_ref{f[3 * i] * d, f[3 * i + 1] * d, f[3 * i + 2] * d};
        jacobian_xyz2uv(xyz_ref, frame_jac);
      }
      float subpix_u_ref = u_ref - u_ref_i;
      float subpix_v_ref = v_ref - v_ref_i;
      float w_ref_tl = (1.0 - subpix_u_ref) * (1.0 - subpix_v_ref);
      float w_ref_tr = subpix_u_ref * (1.0 - subpix_v_ref);
      float w_ref_bl = (1.0 - subpix_u_ref) * subpix_v_ref;
      float w_ref_br = 1.0 - (w_ref_tl + w_ref_tr + w_ref_bl);
      float* cache_ptr = ref_patch_cache.data() + (size_t)PATCH_AREA * i;
      for (int n = 0; n < PATCH_AREA; ++n, ++cache_ptr) {
        const uint8_t* p = ref_img + (v_ref_i + kPattern[offset_all][n][1]) * stride + u_ref_i + kPattern[offset_all][n][0];
        *cache_ptr = w_ref_tl * p[0] + w_ref_tr * p[1] + w_ref_bl * p[stride] + w_ref_br * p[stride + 1];
        if (inverse_comp) {
          float dx = 0.5f * ((w_ref_tl * p[1] + w_ref_tr * p[2] + w_ref_bl * p[stride + 1] + w_ref_br * p[stride + 2]) -
                             (w_ref_tl * p[-1] + w_ref_tr * p[0] + w_ref_bl * p[stride - 1] + w_ref_br * p[stride]));
          float dy = 0.5f * ((w_ref_tl * p[stride] + w_ref_tr * p[1 + stride] + w_ref_bl * p[stride * 2] + w_ref_br * p[stride * 2 + 1]) -
                             (w_ref_tl * p[-stride] + w_ref_tr * p[1 - stride] + w_ref_bl * p[0] + w_ref_br * p[1]));
          double* col = jac_raw.data() + (size_t)6 * ((size_t)i * PATCH_AREA + n);
          for (int k = 0; k < 6; ++k) col[k] = dx * frame_jac[0][k] * fxl + dy * frame_jac[1][k] * fyl;
        }
      }
    }
  }
  struct Proj { bool ok; int ui, vi; float wtl, wtr, wbl, wbr; V3 xyz_cur; };
  // shared head of computeResiduals / selectRobustFunctionLevel (CoarseTracker.cpp:290-323, :557-583)
  Proj project(int i, const SE3& T, float scale, int border) const {
    Proj pr{};
    pr.ok = false;
    double d = dist[i];
    if (d < 0) return pr;
    V3 xyz_ref{f[3 * i] * d, f[3 * i + 1] * d, f[3 * i + 2] * d};
    V3 xyz_cur = T.apply(xyz_ref);
    if (xyz_cur.z < 0) return pr;
    double xyz[3] = {xyz_cur.x, xyz_cur.y, xyz_cur.z}, pxd[2];
    orc_world2cam(cam, xyz, pxd);
    float u0 = (float)pxd[0], v0 = (float)pxd[1];
    float u_cur = u0 * scale, v_cur = v0 * scale;
    int u_cur_i = floorf(u_cur), v_cur_i = floorf(v_cur);
    if (u_cur_i - border < 0 || v_cur_i - border < 0 || u_cur_i + border >= cols || v_cur_i + border >= rows) return pr;
    float su = u_cur - u_cur_i, sv = v_cur - v_cur_i;
    pr.wtl = (1.0 - su) * (1.0 - sv);
    pr.wtr = su * (1.0 - sv);
    pr.wbl = (1.0 - su) * sv;
    pr.wbr = su * sv;
    pr.ui = u_cur_i; pr.vi = v_cur_i; pr.xyz_cur = xyz_cur; pr.ok = true;
    return pr;
  }

  // CoarseTracker.cpp:242-414
  double computeResiduals(const SE3& T_cur_ref, float exposure_rat, double cutoff_error, float b = 0) {
    const int stride = cols;
    const int border = HALF_PATCH_SIZE + 1;
    const float scale = 1.0f / (1 << level);
    const double fxl = cam->fx * scale;
    const double fyl = cam->fy * scale;
    float setting_huberTH = huber_thresh;
    const float max_energy = 2 * setting_huberTH * cutoff_error - setting_huberTH * setting_huberTH;
    buf.clear();
    total_terms = saturated_terms = 0;
    float E = 0;
    for (int i = 0; i < F; ++i) {
      if (!visible[i]) continue;
      Proj pr = project(i, T_cur_ref, scale, border);
      if (!pr.ok) continue;
      double frame_jac[2][6] = {};
      if (!inverse_comp) jacobian_xyz2uv(pr.xyz_cur, frame_jac);
      const float w_cur_tl = pr.wtl, w_cur_tr = pr.wtr, w_cur_bl = pr.wbl, w_cur_br = pr.wbr;
      const float* ref_patch_cache_ptr = ref_patch_cache.data() + (size_t)PATCH_AREA * i;
      for (int n = 0; n < PATCH_AREA; ++n, ++ref_patch_cache_ptr) {
        const uint8_t* p = cur_img + (pr.vi + kPattern[offset_all][n][1]) * stride + pr.ui + kPattern[offset_all][n][0];
        float cur_color = w_cur_tl * p[0] + w_cur_tr * p[1] + w_cur_bl * p[stride] + w_cur_br * p[stride + 1];
        if (!std::isfinite(cur_color)) continue;
        float residual = cur_color - (exposure_rat * (*ref_patch_cache_ptr) + b);
        float hw = std::fabs(residual) < setting_huberTH ? 1 : setting_huberTH / std::fabs(residual);
        if (std::fabs(residual) > cutoff_error && level < max_level) {
          E += max_energy;
          total_terms++;
          saturated_terms++;
        } else {
          if (level == max_level) E += hw * residual * residual;
          else E += hw * residual * residual * (2 - hw);
          total_terms++;
          Term t;
          if (!inverse_comp) {
            float dx = 0.5f * ((w_cur_tl * p[1] + w_cur_tr * p[2] + w_cur_bl * p[stride + 1] + w_cur_br * p[stride + 2]) -
                               (w_cur_tl * p[-1] + w_cur_tr * p[0] + w_cur_bl * p[stride - 1] + w_cur_br * p[stride]));
            float dy = 0.5f * ((w_cur_tl * p[stride] + w_cur_tr * p[1 + stride] + w_cur_bl * p[stride * 2] + w_cur_br * p[stride * 2 + 1]) -
                               (w_cur_tl * p[-stride] + w_cur_tr * p[1 - stride] + w_cur_bl * p[0] + w_cur_br * p[1]));
            for (int k = 0; k < 6; ++k) t.J[1 + k] = dx * frame_jac[0][k] * fxl + dy * frame_jac[1][k] * fyl;
          } else {
            // m_jacobian_cache_true = exposure_rat * m_jacobian_cache_raw (CoarseTracker.cpp:244-245, float * double)
            const double* col = jac_raw.data() + (size_t)6 * ((size_t)i * PATCH_AREA + n);
            for (int k = 0; k < 6; ++k) t.J[1 + k] = exposure_rat * col[k];
          }
          t.J[0] = -(*ref_patch_cache_ptr);
          t.w = hw;
          t.r = residual;
          buf.push_back(t);
        }
      }
    }
    return E / total_terms;
  }

  // CoarseTracker.cpp:499-525
  void computeGS(double H[49], double b[7]) {
    acc.initialize();
    for (int k = 0; k < 7; ++k) b[k] = 0;
    for (const Term& t : buf) {
      acc.update((float)t.J[0], (float)t.J[1], (float)t.J[2], (float)t.J[3], (float)t.J[4], (float)t.J[5], (float)t.J[6], (float)t.w);
      for (int k = 0; k < 7; ++k) b[k] -= t.J[k] * t.r * t.w;
    }
    acc.finish(H);
  }

  // CoarseTracker.cpp:530-644
  int selectRobustFunctionLevel(const SE3& T_cur_ref, float exposure_rat, float b = 0) {
    const int stride = cols;
    const int border = HALF_PATCH_SIZE + 1;
    const float scale = 1.0f / (1 << level);
    std::vector<float> errors;
    for (int i = 0; i < F; ++i) {
      if (!visible[i]) continue;
      Proj pr = project(i, T_cur_ref, scale, border);
      if (!pr.ok) continue;
      const float* ref_patch_cache_ptr = ref_patch_cache.data() + (size_t)PATCH_AREA * i;
      for (int n = 0; n < PATCH_AREA; ++n, ++ref_patch_cache_ptr) {
        const uint8_t* p = cur_img + (pr.vi + kPattern[offset_all][n][1]) * stride + pr.ui + kPattern[offset_all][n][0];
        float cur_color = pr.wtl * p[0] + pr.wtr * p[1] + pr.wbl * p[stride] + pr.wbr * p[stride + 1];
        float residual = cur_color - (exposure_rat * (*ref_patch_cache_ptr) + b);
        errors.push_back(fabsf(residual));
      }
    }
    const int n_err = (int)errors.size();
    if (errors.size() < 30) {
      huber_thresh = 5.2;
      outlier_thresh = 100;
      return n_err;
    }
    float residual_median = median_inplace(errors);
    std::vector<float> absolute_deviation;
    for (size_t i = 0; i < errors.size(); ++i) absolute_deviation.push_back(std::fabs(errors[i] - residual_median));
    float standard_deviation = 1.4826 * median_inplace(absolute_deviation);
    huber_thresh = residual_median + standard_deviation;
    outlier_thresh = 3 * huber_thresh;
    if (outlier_thresh < 10) outlier_thresh = 10;
    return n_err;
  }
};

// CoarseTracker.cpp:112-124
void solve_step(const double H[49], const double b[7], float lambda, double step[7]) {
  double Hl[49];
  for (int i = 0; i < 49; ++i) Hl[i] = H[i];
  for (int i = 0; i < 7; i++) Hl[i * 7 + i] *= (1 + lambda);
  ldlt_solve<7>(Hl, b, step);
  float extrap_fac = 1;
  if (lambda < 0.001) extrap_fac = sqrt(sqrt(0.001 / lambda));
  for (int i = 0; i < 7; ++i) step[i] *= extrap_fac;
  double s = 0;
  for (int i = 0; i < 7; ++i) s += step[i];
  if (!std::isfinite(s) || std::isnan(step[0])) for (int i = 0; i < 7; ++i) step[i] = 0;
}

void fill_trace(orc_trace* e, const Tracker& tr, int iter, const SE3& T, float a, float lambda, const double* H, const double* b,
                const double* step, double energy, int accepted) {
  e->level = tr.level; e->iter = iter;
  T.to_rt(e->T_eval);
  e->a_eval = a; e->lambda = lambda;
  for (int i = 0; i < 49; ++i) e->H[i] = H[i];
  for (int i = 0; i < 7; ++i) { e->b[i] = b[i]; e->step[i] = step ? step[i] : 0.0; }
  e->energy = energy; e->total_terms = tr.total_terms; e->saturated_terms = tr.saturated_terms;
  e->accepted = accepted; e->huber = tr.huber_thresh; e->outlier = tr.outlier_thresh;
}

}  // namespace

extern "C" {

// The tiered float accumulator alone, driven like computeGS drives it (src/CoarseTracker.cpp:507-518): for pinning against the reference's own
// Accumulator7 (oracle/_ref).
void orc_accumulator7(int n, const float* J /*7n*/, const float* w /*n*/, float* H49) {
  Acc7 acc;
  acc.initialize();
  for (int i = 0; i < n; ++i) acc.update(J[7 * i], J[7 * i + 1], J[7 * i + 2], J[7 * i + 3], J[7 * i + 4], J[7 * i + 5], J[7 * i + 6], w[i]);
  double H[49];
  acc.finish(H);
  for (int k = 0; k < 49; ++k) H49[k] = (float)H[k];
}

void orc_make_depth_ref(const double T_ref_w[12], int F, const uint8_t* has_point, const double* f_host, const double* idist,
                        const double* T_host_w, double* dist_out) {
  SE3 Tref = SE3::from_rt(T_ref_w);
  for (int i = 0; i < F; ++i) {
    dist_out[i] = -1;
    if (!has_point[i]) continue;
    double inv = 1.0 / idist[i];
    V3 p_host{f_host[3 * i] * inv, f_host[3 * i + 1] * inv, f_host[3 * i + 2] * inv};
    SE3 T_r_h = Tref.mul(SE3::from_rt(T_host_w + 12 * i).inverse());
    V3 p_ref = T_r_h.apply(p_host);
    if (p_ref.z < 0.00001) continue;
    dist_out[i] = p_ref.norm();
  }
}

uint64_t orc_coarse_track(const orc_cam* cam, const orc_track_params* prm, int n_levels, const uint8_t* const* ref_levels,
                          const uint8_t* const* cur_levels, const int* lw, const int* lh, int F, const double* px, const double* f,
                          const double* dist, double T_cur_ref_io[12], float* a_io, orc_trace* trace, int trace_cap, int* trace_len,
                          int* n_evals_out) {
  int tl = 0, n_evals = 0;
  if (trace_len) *trace_len = 0;
  if (n_evals_out) *n_evals_out = 0;
  if (F == 0) return 0;  // CoarseTracker.cpp:53
  (void)n_levels;
  Tracker tr;
  tr.cam = cam; tr.inverse_comp = prm->inverse_comp != 0; tr.max_level = prm->max_level; tr.min_level = prm->min_level;
  tr.n_iter = prm->n_iter; tr.F = F; tr.px = px; tr.f = f; tr.dist = dist;
  float m_exposure_rat = *a_io;
  SE3 m_T_cur_ref = SE3::from_rt(T_cur_ref_io);
  for (int level = tr.max_level; level >= tr.min_level; --level) {
    tr.set_level(level, ref_levels[level], cur_levels[level], lw[level], lh[level]);
    tr.precomputeReferencePatches();
    tr.selectRobustFunctionLevel(m_T_cur_ref, m_exposure_rat);
    const double cutoff_error = tr.outlier_thresh;
    double energy_old = tr.computeResiduals(m_T_cur_ref, m_exposure_rat, cutoff_error);
    ++n_evals;
    double H[49], b[7];
    tr.computeGS(H, b);
    if (trace && tl < trace_cap) fill_trace(&trace[tl++], tr, -1, m_T_cur_ref, m_exposure_rat, 0.f, H, b, nullptr, energy_old, 1);
    float lambda = 0.1;
    for (int iter = 0; iter < tr.n_iter; iter++) {
      double step[7];
      solve_step(H, b, lambda, step);
      float new_exposure_rat = m_exposure_rat + step[0];
      double neg[6];
      for (int k = 0; k < 6; ++k) neg[k] = -step[1 + k];
      SE3 new_T_cur_ref = !tr.inverse_comp ? SE3::exp(neg).mul(m_T_cur_ref) : m_T_cur_ref.mul(SE3::exp(neg));
      double energy_new = tr.computeResiduals(new_T_cur_ref, new_exposure_rat, cutoff_error);
      ++n_evals;
      const bool accepted = energy_new < energy_old;
      if (trace && tl < trace_cap) fill_trace(&trace[tl++], tr, iter, new_T_cur_ref, new_exposure_rat, lambda, H, b, step, energy_new, accepted);
      if (accepted) {
        tr.computeGS(H, b);
        energy_old = energy_new;
        m_exposure_rat = new_exposure_rat;
        m_T_cur_ref = new_T_cur_ref;
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < 0.001) lambda = 0.001;
      }
      double nrm = 0;
      for (int k = 0; k < 7; ++k) nrm += step[k] * step[k];
      nrm = std::sqrt(nrm);
      if (!(nrm > 1e-4)) break;
    }
  }
  m_T_cur_ref.to_rt(T_cur_ref_io);
  *a_io = m_exposure_rat;
  if (trace_len) *trace_len = tl;
  if (n_evals_out) *n_evals_out = n_evals;
  return (uint64_t)(float(tr.total_terms) / tr.PATCH_AREA);  // CoarseTracker.cpp:207
}

void orc_track_eval(const orc_cam* cam, int inverse_comp, int level, int max_level, const uint8_t* ref_img, const uint8_t* cur_img, int w,
                    int h, int F, const double* px, const double* f, const double* dist, const double T[12], float a, float huber,
                    float outlier, double H_out[49], double b_out[7], double* energy_out, int* total_terms, int* saturated_terms) {
  Tracker tr;
  tr.cam = cam; tr.inverse_comp = inverse_comp != 0; tr.max_level = max_level; tr.min_level = level; tr.n_iter = 0;
  tr.F = F; tr.px = px; tr.f = f; tr.dist = dist;
  tr.set_level(level, ref_img, cur_img, w, h);
  tr.precomputeReferencePatches();
  tr.huber_thresh = huber; tr.outlier_thresh = outlier;
  double E = tr.computeResiduals(SE3::from_rt(T), a, (double)outlier);
  tr.computeGS(H_out, b_out);
  *energy_out = E; *total_terms = tr.total_terms; *saturated_terms = tr.saturated_terms;
}

void orc_track_select_robust(const orc_cam* cam, int level, int max_level, const uint8_t* ref_img, const uint8_t* cur_img, int w, int h,
                             int F, const double* px, const double* f, const double* dist, const double T[12], float a, float* huber_out,
                             float* outlier_out, int* n_errors) {
  Tracker tr;
  tr.cam = cam; tr.inverse_comp = false; tr.max_level = max_level; tr.min_level = level; tr.n_iter = 0;
  tr.F = F; tr.px = px; tr.f = f; tr.dist = dist;
  tr.set_level(level, ref_img, cur_img, w, h);
  tr.precomputeReferencePatches();
  int n = tr.selectRobustFunctionLevel(SE3::from_rt(T), a);
  *huber_out = tr.huber_thresh; *outlier_out = tr.outlier_thresh;
  if (n_errors) *n_errors = n;
}

void orc_track_solve(const double H[49], const double b[7], float lambda, double step_out[7]) { solve_step(H, b, lambda, step_out); }

}  // extern "C"
