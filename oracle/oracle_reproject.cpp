// ORACLE — test infrastructure, never linked into or called by the product (hso_b200/).
// CPU restatement of row N1: the data path of Reprojector::reprojectMap (src/reprojector.cpp:88-331) on flattened inputs:
// reprojectPoint (:504-529), the grid cells as std::list with list::sort(pointQualityComparator) (:333-344,355-356), the three
// selection passes over grid_.cell_order (:262-303), reprojectCellAll (:545-615), and the whole Matcher::findMatchDirect
// (src/matcher.cpp:270-375) with warp::getWarpMatrixAffine (:46-72), getBestSearchLevel (:74-85), cam2world
// (src/camera.cpp:66-87 pinhole incl. the cv::undistortPoints branch, :169-190 FOV, :297-300 equidistant).
//
// cv::undistortPoints is OpenCV (not vendored, unpinned): restated from the published algorithm (modules/calib3d/src/undistort.dispatch.cpp,
// cvUndistortPointsInternal: 5 fixed-point iterations, float in/out, float camera matrix as the reference builds it, camera.cpp:43-45)
// and pinned against cv2 4.13 golden vectors (tests/golden/cv_golden2.npz).
#include <cmath>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

#include "hso_oracle.h"
#include "oracle_math.hpp"

using namespace orc;

extern "C" {

// cv::undistortPoints for one CV_32FC2 point with a float camera matrix K = (fx, fy, cx, cy) and float distortion (k1, k2, p1, p2, k3), no R / P:
// cvUndistortPointsInternal's 5 fixed-point iterations in double, float in / float out.
void orc_cv_undistort_point(const float K[4], const float D[5], float uf, float vf, float out[2]) {
  const double fx = (double)K[0], fy = (double)K[1], cx = (double)K[2], cy = (double)K[3];
  double k[5];
  for (int i = 0; i < 5; ++i) k[i] = (double)D[i];
  const double ifx = 1. / fx, ify = 1. / fy;
  double xx = ((double)uf - cx) * ifx, yy = ((double)vf - cy) * ify;
  const double x0 = xx, y0 = yy;
  for (int j = 0; j < 5; ++j) {
    const double r2 = xx * xx + yy * yy;
    const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) { xx = ((double)uf - cx) * ifx; yy = ((double)vf - cy) * ify; break; }
    const double deltaX = 2 * k[2] * xx * yy + k[3] * (r2 + 2 * xx * xx);
    const double deltaY = k[2] * (r2 + 2 * yy * yy) + 2 * k[3] * xx * yy;
    xx = (x0 - deltaX) * icdist;
    yy = (y0 - deltaY) * icdist;
  }
  out[0] = (float)xx;
  out[1] = (float)yy;
}

// src/camera.cpp:66-87 (PinholeCamera), :169-190 (FOVCamera), :297-300 (EquidistantCamera)
void orc_cam2world(const orc_cam* cam, double u, double v, double xyz_out[3]) {
  double x, y;
  const bool distortion = cam->model == 0 && std::fabs(cam->d[0]) > 0.0000001;  // camera.cpp:36
  if (cam->model == 0 && distortion) {
    // cv::undistortPoints(src(1x1 CV_32FC2), dst, cvK_ (float 3x3), cvD_ (float 1x5)) — no R, no P, criteria (MAX_ITER, 5)
    const float K[4] = {(float)cam->fx, (float)cam->fy, (float)cam->cx, (float)cam->cy};
    float D[5], out[2];
    for (int i = 0; i < 5; ++i) D[i] = (float)cam->d[i];
    orc_cv_undistort_point(K, D, (float)u, (float)v, out);
    x = (double)out[0];
    y = (double)out[1];
  } else if (cam->model == 1 && !cam->undistort) {
    const double ud = (u - cam->cx) / cam->fx, vd = (v - cam->cy) / cam->fy;
    const double dist = std::sqrt(ud * ud + vd * vd);
    const double omega = cam->d[0];
    const double radial_distortion = std::tan(dist * omega) / (2 * dist * std::tan(omega / 2));
    x = radial_distortion * ud;
    y = radial_distortion * vd;
  } else {
    x = (u - cam->cx) / cam->fx;
    y = (v - cam->cy) / cam->fy;
  }
  const double n = std::sqrt(x * x + y * y + 1.0);  // Vector3d::normalized()
  xyz_out[0] = x / n; xyz_out[1] = y / n; xyz_out[2] = 1.0 / n;
}

// warp::getWarpMatrixAffine — src/matcher.cpp:46-72
void orc_get_warp_matrix_affine(const orc_cam* cam, const double px_ref[2], const double f_ref[3], double depth_ref, const double T_cur_ref[12],
                                int level_ref, double A_cur_ref[4]) {
  const int halfpatch_size = 5;
  const SE3 T = SE3::from_rt(T_cur_ref);
  const V3 xyz_ref{f_ref[0] * depth_ref, f_ref[1] * depth_ref, f_ref[2] * depth_ref};
  const int ratio = (1 << level_ref);
  double du[3], dv[3];
  orc_cam2world(cam, px_ref[0] + (double)(halfpatch_size * ratio), px_ref[1], du);
  orc_cam2world(cam, px_ref[0], px_ref[1] + (double)(halfpatch_size * ratio), dv);
  const double sdu = xyz_ref.z / du[2], sdv = xyz_ref.z / dv[2];
  const V3 xyz_du{du[0] * sdu, du[1] * sdu, du[2] * sdu}, xyz_dv{dv[0] * sdv, dv[1] * sdv, dv[2] * sdv};
  double pc[2], pu[2], pv[2];
  const V3 a = T.apply(xyz_ref), b = T.apply(xyz_du), c = T.apply(xyz_dv);
  const double av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z}, cv[3] = {c.x, c.y, c.z};
  orc_world2cam(cam, av, pc);
  orc_world2cam(cam, bv, pu);
  orc_world2cam(cam, cv, pv);
  A_cur_ref[0] = (pu[0] - pc[0]) / halfpatch_size; A_cur_ref[2] = (pu[1] - pc[1]) / halfpatch_size;  // col(0)
  A_cur_ref[1] = (pv[0] - pc[0]) / halfpatch_size; A_cur_ref[3] = (pv[1] - pc[1]) / halfpatch_size;  // col(1)
}

namespace {

struct Cand { int idx; double px[2]; };

struct MatchCtx {
  const orc_cam* cam;
  SE3 T_cur_w;
  std::vector<SE3> T_f_w;
  const orc_reproj_cand* cands;
  const uint8_t* const* const* ref_levels;  // [frame][level]
  const uint8_t* const* cur_levels;
  const int* lw; const int* lh;
  const int16_t* const* cur_sobx; const int16_t* const* cur_soby;
  int align_max_iter, max_search_level;
  orc_reproj_result* out;
};

// Matcher::findMatchDirect — src/matcher.cpp:270-375
bool find_match_direct(MatchCtx& m, int idx, double px_cur[2]) {
  const orc_reproj_cand& c = m.cands[idx];
  orc_reproj_result& r = m.out[idx];
  if (c.ref_pose < 0) return false;  // !pt.getCloseViewObs(...)
  // isInFrame((px/(1<<level)).cast<int>(), halfpatch_size_+2, level) — camera.h:85-89 (integer width()/(1<<level))
  {
    const int lv = c.ref_level;
    const int ox = (int)(c.px_ref[0] / (1 << lv)), oy = (int)(c.px_ref[1] / (1 << lv));
    const int boundary = 4 + 2;
    if (!(ox >= boundary && ox < m.cam->width / (1 << lv) - boundary && oy >= boundary && oy < m.cam->height / (1 << lv) - boundary)) return false;
  }
  const SE3 T_c_r = m.T_cur_w.mul(m.T_f_w[c.ref_pose].inverse());
  double rt[12];
  T_c_r.to_rt(rt);
  double A[4];
  {
    // same arithmetic as orc_get_warp_matrix_affine, on the SE3 itself (no rt round trip)
    const int halfpatch_size = 5;
    const V3 xyz_ref{c.f_ref[0] * c.depth_ref, c.f_ref[1] * c.depth_ref, c.f_ref[2] * c.depth_ref};
    const int ratio = (1 << c.ref_level);
    double du[3], dv[3];
    orc_cam2world(m.cam, c.px_ref[0] + (double)(halfpatch_size * ratio), c.px_ref[1], du);
    orc_cam2world(m.cam, c.px_ref[0], c.px_ref[1] + (double)(halfpatch_size * ratio), dv);
    const double sdu = xyz_ref.z / du[2], sdv = xyz_ref.z / dv[2];
    const V3 xyz_du{du[0] * sdu, du[1] * sdu, du[2] * sdu}, xyz_dv{dv[0] * sdv, dv[1] * sdv, dv[2] * sdv};
    const V3 a = T_c_r.apply(xyz_ref), b = T_c_r.apply(xyz_du), cc = T_c_r.apply(xyz_dv);
    const double av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z}, cv[3] = {cc.x, cc.y, cc.z};
    double pc[2], pu[2], pv[2];
    orc_world2cam(m.cam, av, pc);
    orc_world2cam(m.cam, bv, pu);
    orc_world2cam(m.cam, cv, pv);
    A[0] = (pu[0] - pc[0]) / halfpatch_size; A[2] = (pu[1] - pc[1]) / halfpatch_size;
    A[1] = (pv[0] - pc[0]) / halfpatch_size; A[3] = (pv[1] - pc[1]) / halfpatch_size;
  }
  const int search_level = orc_get_best_search_level(A, m.max_search_level);
  for (int k = 0; k < 4; ++k) r.A_cur_ref[k] = A[k];
  r.search_level = search_level;
  orc_align_job job;
  std::memset(&job, 0, sizeof job);
  job.ref_level = c.ref_level; job.search_level = search_level; job.type = c.ftr_type; job.scale_patch = c.scale_patch;
  job.px_ref[0] = c.px_ref[0]; job.px_ref[1] = c.px_ref[1];
  for (int k = 0; k < 4; ++k) job.A_cur_ref[k] = A[k];
  job.grad[0] = c.grad[0]; job.grad[1] = c.grad[1];
  job.px_cur[0] = px_cur[0]; job.px_cur[1] = px_cur[1];
  job.exposure_rat = c.exposure_rat;
  orc_align_result res;
  orc_match_direct_batch(1, &job, m.ref_levels[c.ref_frame], m.cur_levels, m.lw, m.lh, m.cur_sobx, m.cur_soby, m.align_max_iter, &res);
  px_cur[0] = res.px_cur[0]; px_cur[1] = res.px_cur[1];
  return res.ok != 0;
}

}  // namespace

}  // extern "C"

namespace {

// Cells, ordering and the selection passes of Reprojector::reprojectMap (src/reprojector.cpp:253-303) / reprojectCellAll (:545-615).
// match(idx, px) plays Matcher::findMatchDirect. out[i].in_frame / cell / px are already filled by the reprojectPoint stage.
template <class Match>
void selection_walk(int M, const orc_reproj_cand* cands, const orc_reproj_grid* grid, const int32_t* cell_order, Match&& match,
                    orc_reproj_result* out, orc_reproj_summary* summary) {
  const int n_cells = grid->n_cols * grid->n_rows;
  std::vector<std::list<Cand>> cells(n_cells);
  std::vector<Cand> all;  // allPixelToDistribute
  for (int i = 0; i < M; ++i) {
    if (!out[i].in_frame) continue;
    Cand cd{i, {out[i].px[0], out[i].px[1]}};
    cells.at(out[i].cell).push_back(cd);
    all.push_back(cd);
  }
  size_t n_matches = 0, n_trials = 0;
  int order = 0;
  const size_t maxFts = (size_t)grid->max_fts;
  auto on_match = [&](Cand& cd) {
    orc_reproj_result& r = out[cd.idx];
    r.matched = 1; r.order = order++;
  };
  // pointQualityComparator — reprojector.cpp:333-344
  auto quality = [&](const Cand& lhs, const Cand& rhs) {
    const orc_reproj_cand& a = cands[lhs.idx];
    const orc_reproj_cand& b = cands[rhs.idx];
    if (a.pt_type != b.pt_type) return a.pt_type > b.pt_type;
    if (a.pt_ftr_type > b.pt_ftr_type) return true;
    return false;
  };
  // Reprojector::reprojectCell — reprojector.cpp:351-424
  auto reproject_cell = [&](std::list<Cand>& cell, bool is_2nd, bool is_3rd) -> bool {
    if (cell.empty()) return false;
    if (!is_2nd) cell.sort(quality);
    auto it = cell.begin();
    int succees = 0;
    while (it != cell.end()) {
      ++n_trials;  // counted before the TYPE_DELETED test, reprojector.cpp:361-367
      if (cands[it->idx].pt_type == 0) { it = cell.erase(it); continue; }
      out[it->idx].tried = 1;
      const bool ok = match(it->idx, it->px);
      out[it->idx].px[0] = it->px[0]; out[it->idx].px[1] = it->px[1];
      if (!ok) { it = cell.erase(it); continue; }
      on_match(*it);
      it = cell.erase(it);
      if (!is_3rd) return true;
      succees++;
      n_matches++;
      if (succees >= 3 || n_matches >= maxFts) return true;
    }
    return false;
  };
  summary->used_cell_all = 0;
  if (all.size() < maxFts + 50) {
    // Reprojector::reprojectCellAll — reprojector.cpp:545-615
    summary->used_cell_all = 1;
    for (auto& cd : all) {
      ++n_trials;  // reprojector.cpp:553-559
      if (cands[cd.idx].pt_type == 0) continue;
      out[cd.idx].tried = 1;
      const bool ok = match(cd.idx, cd.px);
      out[cd.idx].px[0] = cd.px[0]; out[cd.idx].px[1] = cd.px[1];
      if (!ok) continue;
      on_match(cd);
      n_matches++;
      if (n_matches >= maxFts) break;
    }
  } else {
    for (size_t i = 0; i < cells.size(); ++i) {  // 1st
      if (reproject_cell(cells.at(cell_order[i]), false, false)) ++n_matches;
      if (n_matches >= maxFts) break;
    }
    if (n_matches < maxFts) {  // 2nd (the reference's loop never visits i == 0)
      for (size_t i = cells.size() - 1; i > 0; --i) {
        if (reproject_cell(cells.at(cell_order[i]), true, false)) ++n_matches;
        if (n_matches >= maxFts) break;
      }
    }
    if (n_matches < maxFts) {  // 3rd
      for (size_t i = 0; i < cells.size(); ++i) {
        reproject_cell(cells.at(cell_order[i]), true, true);
        if (n_matches >= maxFts) break;
      }
    }
  }
  int n_in = 0;
  for (int i = 0; i < M; ++i) n_in += out[i].in_frame ? 1 : 0;
  summary->n_in_frame = n_in;
  summary->n_matches = (int)n_matches;
  summary->n_trials = (int)n_trials;
}

}  // namespace

extern "C" {

void orc_reproject_select(int M, const orc_reproj_cand* cands, const uint8_t* match_ok, const orc_reproj_grid* grid, const int32_t* cell_order,
                          orc_reproj_result* io, orc_reproj_summary* summary) {
  for (int i = 0; i < M; ++i) { io[i].tried = 0; io[i].matched = 0; io[i].order = -1; }
  selection_walk(M, cands, grid, cell_order, [&](int idx, double*) { return match_ok[idx] != 0; }, io, summary);
}

// Context + Reprojector::reprojectPoint (src/reprojector.cpp:504-529) for every candidate: fills in_frame / cell / px.
static void reproject_points(MatchCtx& m, const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int M,
                             const orc_reproj_cand* cands, const orc_reproj_grid* grid, int max_search_level,
                             const uint8_t* const* const* ref_levels, const uint8_t* const* cur_levels, const int* lw, const int* lh,
                             const int16_t* const* cur_sobx, const int16_t* const* cur_soby, orc_reproj_result* out) {
  m.cam = cam;
  m.T_cur_w = SE3::from_rt(T_cur_w);
  for (int k = 0; k < n_poses; ++k) m.T_f_w.push_back(SE3::from_rt(T_f_w + 12 * k));
  m.cands = cands; m.ref_levels = ref_levels; m.cur_levels = cur_levels; m.lw = lw; m.lh = lh; m.cur_sobx = cur_sobx; m.cur_soby = cur_soby;
  m.align_max_iter = grid->align_max_iter; m.max_search_level = max_search_level; m.out = out;
  for (int i = 0; i < M; ++i) {
    orc_reproj_result& r = out[i];
    std::memset(&r, 0, sizeof r);
    r.order = -1; r.cell = -1;
    const orc_reproj_cand& c = cands[i];
    const V3 pHost{c.p_host[0], c.p_host[1], c.p_host[2]};
    const V3 pTarget = m.T_cur_w.mul(m.T_f_w[c.host_pose].inverse()).apply(pHost);
    if (pTarget.z < 0.00001) continue;
    const double pt[3] = {pTarget.x, pTarget.y, pTarget.z};
    double px[2];
    orc_world2cam(cam, pt, px);
    r.px[0] = px[0]; r.px[1] = px[1];
    const int ox = (int)px[0], oy = (int)px[1];  // px.cast<int>()
    if (ox >= 8 && ox < cam->width - 8 && oy >= 8 && oy < cam->height - 8) {  // isInFrame(px.cast<int>(), 8)
      r.in_frame = 1;
      r.cell = static_cast<int>(px[1] / grid->cell_size) * grid->n_cols + static_cast<int>(px[0] / grid->cell_size);
    }
  }
}

void orc_reproject_match(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int M, const orc_reproj_cand* cands,
                         const orc_reproj_grid* grid, const int32_t* cell_order, int max_search_level,
                         const uint8_t* const* const* ref_levels /*[frame][level]*/, const uint8_t* const* cur_levels, const int* lw, const int* lh,
                         const int16_t* const* cur_sobx, const int16_t* const* cur_soby, orc_reproj_result* out, orc_reproj_summary* summary) {
  MatchCtx m;
  reproject_points(m, cam, T_cur_w, n_poses, T_f_w, M, cands, grid, max_search_level, ref_levels, cur_levels, lw, lh, cur_sobx, cur_soby, out);
  selection_walk(M, cands, grid, cell_order, [&](int idx, double* px) {
    const bool ok = find_match_direct(m, idx, px);
    out[idx].align_ok = ok ? 1 : 0;
    return ok;
  }, out, summary);
}

// findMatchDirect evaluated for EVERY candidate that entered a cell (what the CUDA path computes speculatively before replaying the
// selection): in_frame / cell / px as above, align_ok / A_cur_ref / search_level / px after the alignment in px_after.
void orc_reproject_speculative(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int M, const orc_reproj_cand* cands,
                               const orc_reproj_grid* grid, int max_search_level, const uint8_t* const* const* ref_levels,
                               const uint8_t* const* cur_levels, const int* lw, const int* lh, const int16_t* const* cur_sobx,
                               const int16_t* const* cur_soby, orc_reproj_result* out, double* px_after /*2M*/) {
  MatchCtx m;
  reproject_points(m, cam, T_cur_w, n_poses, T_f_w, M, cands, grid, max_search_level, ref_levels, cur_levels, lw, lh, cur_sobx, cur_soby, out);
  for (int i = 0; i < M; ++i) {
    px_after[2 * i] = out[i].px[0]; px_after[2 * i + 1] = out[i].px[1];
    if (!out[i].in_frame || cands[i].pt_type == 0) continue;
    double px[2] = {out[i].px[0], out[i].px[1]};
    out[i].align_ok = find_match_direct(m, i, px) ? 1 : 0;
    px_after[2 * i] = px[0]; px_after[2 * i + 1] = px[1];
  }
}

}  // extern "C"

// ---- a13b: seed stage of Reprojector::reprojectMap -------------------------------------------------------------------------------------------
namespace {

struct SeedCtx {
  const orc_cam* cam;
  SE3 T_cur_w;
  std::vector<SE3> T_f_w;
  const orc_seed_obs* seeds;
  const uint8_t* const* const* ref_levels;
  const uint8_t* const* cur_levels;
  const int* lw; const int* lh;
  const int16_t* const* cur_sobx; const int16_t* const* cur_soby;
  int align_max_iter, max_search_level;
  orc_reproj_result* out;
};

// Matcher::findMatchSeed — src/matcher.cpp:442-518
bool find_match_seed(SeedCtx& m, int idx, double px_cur[2]) {
  const orc_seed_obs& s = m.seeds[idx];
  orc_reproj_result& r = m.out[idx];
  const SE3& T_ref = m.T_f_w[s.ref_pose];
  const SE3 T_ref_inv = T_ref.inverse();
  const double inv_mu = 1.0 / s.mu;  // 1.0/seed.mu: float promoted to double
  // compute parallax angle
  const V3 seed_pos = T_ref_inv.apply(V3{inv_mu * s.f[0], inv_mu * s.f[1], inv_mu * s.f[2]});
  V3 ref_dir = T_ref_inv.t - seed_pos;            // Frame::pos() = T_f_w_.inverse().translation()
  { const double n = ref_dir.norm(); ref_dir = V3{ref_dir.x / n, ref_dir.y / n, ref_dir.z / n}; }  // Eigen normalize(): divides by the norm
  V3 cur_dir = m.T_cur_w.inverse().t - seed_pos;
  { const double n = cur_dir.norm(); cur_dir = V3{cur_dir.x / n, cur_dir.y / n, cur_dir.z / n}; }
  const double cos_angle = ref_dir.dot(cur_dir);
  if (cos_angle < 0.5) return false;
  {
    const int lv = s.level;
    const int ox = (int)(s.px[0] / (1 << lv)), oy = (int)(s.px[1] / (1 << lv));
    const int boundary = 4 + 2;
    if (!(ox >= boundary && ox < m.cam->width / (1 << lv) - boundary && oy >= boundary && oy < m.cam->height / (1 << lv) - boundary)) return false;
  }
  const SE3 T_c_r = m.T_cur_w.mul(T_ref_inv);
  double A[4];
  {
    const int halfpatch_size = 5;
    const double depth_ref = 1. / s.mu;
    const V3 xyz_ref{s.f[0] * depth_ref, s.f[1] * depth_ref, s.f[2] * depth_ref};
    const int ratio = (1 << s.level);
    double du[3], dv[3];
    orc_cam2world(m.cam, s.px[0] + (double)(halfpatch_size * ratio), s.px[1], du);
    orc_cam2world(m.cam, s.px[0], s.px[1] + (double)(halfpatch_size * ratio), dv);
    const double sdu = xyz_ref.z / du[2], sdv = xyz_ref.z / dv[2];
    const V3 xyz_du{du[0] * sdu, du[1] * sdu, du[2] * sdu}, xyz_dv{dv[0] * sdv, dv[1] * sdv, dv[2] * sdv};
    const V3 a = T_c_r.apply(xyz_ref), b = T_c_r.apply(xyz_du), cc = T_c_r.apply(xyz_dv);
    const double av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z}, cv[3] = {cc.x, cc.y, cc.z};
    double pc[2], pu[2], pv[2];
    orc_world2cam(m.cam, av, pc);
    orc_world2cam(m.cam, bv, pu);
    orc_world2cam(m.cam, cv, pv);
    A[0] = (pu[0] - pc[0]) / halfpatch_size; A[2] = (pu[1] - pc[1]) / halfpatch_size;
    A[1] = (pv[0] - pc[0]) / halfpatch_size; A[3] = (pv[1] - pc[1]) / halfpatch_size;
  }
  const int search_level = orc_get_best_search_level(A, m.max_search_level);
  for (int k = 0; k < 4; ++k) r.A_cur_ref[k] = A[k];
  r.search_level = search_level;
  orc_align_job job;
  std::memset(&job, 0, sizeof job);
  job.ref_level = s.level; job.search_level = search_level; job.type = s.ftr_type;
  job.scale_patch = std::fabs(s.exposure_rat * 128 - 128) > 30.0f ? 1 : 0;  // fabsf(exposure_rat*128 - 128) > LIGHT_THRESHOLD (:473), no keyframe-gap test
  job.px_ref[0] = s.px[0]; job.px_ref[1] = s.px[1];
  for (int k = 0; k < 4; ++k) job.A_cur_ref[k] = A[k];
  job.grad[0] = s.grad[0]; job.grad[1] = s.grad[1];
  job.px_cur[0] = px_cur[0]; job.px_cur[1] = px_cur[1];
  job.exposure_rat = s.exposure_rat;
  job.ncc_thresh = 0.8f;  // checkNCC(patch_f_, patchNCC, 0.8) (:510); the ncc_thresh argument of findMatchSeed is unused
  orc_align_result res;
  orc_match_direct_batch(1, &job, m.ref_levels[s.ref_frame], m.cur_levels, m.lw, m.lh, m.cur_sobx, m.cur_soby, m.align_max_iter, &res);
  px_cur[0] = res.px_cur[0]; px_cur[1] = res.px_cur[1];
  return res.ok != 0;
}

// Reprojector::reprojectorSeed — src/reprojector.cpp:531-552
void reproject_seed_points(SeedCtx& m, const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int S, const orc_seed_obs* seeds,
                           const orc_reproj_grid* grid, int max_search_level, const uint8_t* const* const* ref_levels, const uint8_t* const* cur_levels,
                           const int* lw, const int* lh, const int16_t* const* cur_sobx, const int16_t* const* cur_soby, orc_reproj_result* out) {
  m.cam = cam;
  m.T_cur_w = SE3::from_rt(T_cur_w);
  for (int k = 0; k < n_poses; ++k) m.T_f_w.push_back(SE3::from_rt(T_f_w + 12 * k));
  m.seeds = seeds; m.ref_levels = ref_levels; m.cur_levels = cur_levels; m.lw = lw; m.lh = lh; m.cur_sobx = cur_sobx; m.cur_soby = cur_soby;
  m.align_max_iter = grid->align_max_iter; m.max_search_level = max_search_level; m.out = out;
  for (int i = 0; i < S; ++i) {
    orc_reproj_result& r = out[i];
    std::memset(&r, 0, sizeof r);
    r.order = -1; r.cell = -1;
    const orc_seed_obs& s = seeds[i];
    const SE3 Tth = m.T_cur_w.mul(m.T_f_w[s.ref_pose].inverse());
    const double inv_mu = 1.0 / s.mu;
    const V3 pTarget = Tth.apply(V3{inv_mu * s.f[0], inv_mu * s.f[1], inv_mu * s.f[2]});
    if (pTarget.z < 0.001) continue;
    const double pt[3] = {pTarget.x, pTarget.y, pTarget.z};
    double px[2];
    orc_world2cam(cam, pt, px);
    r.px[0] = px[0]; r.px[1] = px[1];
    const int ox = (int)px[0], oy = (int)px[1];
    if (ox >= 8 && ox < cam->width - 8 && oy >= 8 && oy < cam->height - 8) {
      r.in_frame = 1;
      r.cell = static_cast<int>(px[1] / grid->cell_size) * grid->n_cols + static_cast<int>(px[0] / grid->cell_size);
    }
  }
}

struct SeedCand { int idx; double px[2]; };

// The seed loop of reprojectMap (src/reprojector.cpp:318-327) + Reprojector::reprojectorSeeds (:431-503); match plays findMatchSeed.
template <class Match>
void seed_walk(int S, const orc_seed_obs* seeds, const orc_reproj_grid* grid, const int32_t* cell_order, int n_matches_in, Match&& match,
               orc_reproj_result* out, orc_reproj_summary* summary) {
  const int n_cells = grid->n_cols * grid->n_rows;
  std::vector<std::list<SeedCand>> cells(n_cells);
  int n_in = 0;
  for (int i = 0; i < S; ++i) {
    if (!out[i].in_frame) continue;
    ++n_in;
    cells.at(out[i].cell).push_back(SeedCand{i, {out[i].px[0], out[i].px[1]}});
  }
  size_t n_matches = (size_t)n_matches_in, n_trials = 0;
  int order = 0;
  const size_t maxFts = (size_t)grid->max_fts;
  auto seed_cmp = [&](const SeedCand& l, const SeedCand& r) { return seeds[l.idx].sigma2 < seeds[r.idx].sigma2; };  // seedComparator :346-349
  auto reproject_seeds = [&](std::list<SeedCand>& sell) -> bool {
    sell.sort(seed_cmp);
    auto it = sell.begin();
    while (it != sell.end()) {
      ++n_trials;
      out[it->idx].tried = 1;
      const bool ok = match(it->idx, it->px);
      out[it->idx].px[0] = it->px[0]; out[it->idx].px[1] = it->px[1];
      if (ok) {
        out[it->idx].matched = 1; out[it->idx].order = order++;
        it = sell.erase(it);
        return true;
      }
      ++it;
    }
    return false;
  };
  for (size_t i = 0; i < cells.size(); ++i) {
    if (reproject_seeds(cells.at(cell_order[i]))) ++n_matches;
    if (n_matches >= maxFts) break;
  }
  summary->n_in_frame = n_in;
  summary->n_matches = (int)n_matches;
  summary->n_trials = (int)n_trials;
  summary->used_cell_all = 0;
}

}  // namespace

extern "C" {

void orc_seed_select(int S, const orc_seed_obs* seeds, const uint8_t* match_ok, const orc_reproj_grid* grid, const int32_t* cell_order, int n_matches_in,
                     orc_reproj_result* io, orc_reproj_summary* summary) {
  for (int i = 0; i < S; ++i) { io[i].tried = 0; io[i].matched = 0; io[i].order = -1; }
  seed_walk(S, seeds, grid, cell_order, n_matches_in, [&](int idx, double*) { return match_ok[idx] != 0; }, io, summary);
}

void orc_reproject_seeds(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int S, const orc_seed_obs* seeds,
                         const orc_reproj_grid* grid, const int32_t* cell_order, int n_matches_in, int max_search_level,
                         const uint8_t* const* const* ref_levels, const uint8_t* const* cur_levels, const int* lw, const int* lh,
                         const int16_t* const* cur_sobx, const int16_t* const* cur_soby, orc_reproj_result* out, orc_reproj_summary* summary) {
  SeedCtx m;
  reproject_seed_points(m, cam, T_cur_w, n_poses, T_f_w, S, seeds, grid, max_search_level, ref_levels, cur_levels, lw, lh, cur_sobx, cur_soby, out);
  seed_walk(S, seeds, grid, cell_order, n_matches_in, [&](int idx, double* px) {
    const bool ok = find_match_seed(m, idx, px);
    out[idx].align_ok = ok ? 1 : 0;
    return ok;
  }, out, summary);
}

void orc_reproject_seeds_speculative(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int S, const orc_seed_obs* seeds,
                                     const orc_reproj_grid* grid, int max_search_level, const uint8_t* const* const* ref_levels,
                                     const uint8_t* const* cur_levels, const int* lw, const int* lh, const int16_t* const* cur_sobx,
                                     const int16_t* const* cur_soby, orc_reproj_result* out, double* px_after) {
  SeedCtx m;
  reproject_seed_points(m, cam, T_cur_w, n_poses, T_f_w, S, seeds, grid, max_search_level, ref_levels, cur_levels, lw, lh, cur_sobx, cur_soby, out);
  for (int i = 0; i < S; ++i) {
    px_after[2 * i] = out[i].px[0]; px_after[2 * i + 1] = out[i].px[1];
    if (!out[i].in_frame) continue;
    double px[2] = {out[i].px[0], out[i].px[1]};
    out[i].align_ok = find_match_seed(m, i, px) ? 1 : 0;
    px_after[2 * i] = px[0]; px_after[2 * i + 1] = px[1];
  }
}

}  // extern "C"
