// C++ host layer above the C-ABI (include/hso_b200.h): the reference's class / function surface for the tracking hot path,
// re-stated on dependency-free types so it builds without Eigen / Sophus / OpenCV / Boost (none are present in this image).
// Inside the reference tree the same calls are made from the real classes; INTEGRATION.md shows that glue line by line.
//
//   reference                                                        here
//   hso::Frame::Frame(cam, img, ts)            frame.h:131          hso::b200::Frame(ctx, img, W, H, stride, ts)   (throws on bad size)
//   hso::CoarseTracker(ic,max,min,n_iter,v)    CoarseTracker.h:134  hso::b200::CoarseTracker(ctx, ic, max, min, n_iter, v)
//   size_t CoarseTracker::run(ref, cur)        CoarseTracker.h:141  size_t run(FramePtr ref, FramePtr cur)
//   CoarseTracker::makeDepthRef()              CoarseTracker.cpp:210  makeDepthRef(ref)  (host: pointer chasing over Feature/Point)
//   pose_optimizer::optimizeLevenbergMarquardt3rd(...)  pose_optimizer.h:61  hso::b200::pose_optimizer::optimizeLevenbergMarquardt3rd(...)
//   Matcher::findMatchDirect(pt, cur, px)      matcher.h:153        Matcher::findMatchDirectBatch(candidates, cur)  (after getWarpMatrixAffine)
//   Reprojector::reprojectMap(frame, ...)      reprojector.h        Reprojector::reprojectMap(frame, points, keyframes)  (row N1: one device call)
//   FrameHandlerMono::addImage front end       frame_handler_mono.cpp:92,190-204  addImagesAndTrack(...)  (B streams, chunk-pipelined)
//
// Poses are 3x4 row-major [R|t] (SE3::matrix3x4()). All arithmetic of the path runs on the GPU; this layer only flattens.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hso_b200.h"

namespace hso {
namespace b200 {

struct SE3 {  // 3x4 row-major [R | t]
  double m[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  SE3 inverse() const {
    SE3 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r.m[4 * i + j] = m[4 * j + i];
    for (int i = 0; i < 3; ++i) r.m[4 * i + 3] = -(r.m[4 * i] * m[3] + r.m[4 * i + 1] * m[7] + r.m[4 * i + 2] * m[11]);
    return r;
  }
  SE3 operator*(const SE3& o) const {
    SE3 r;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) r.m[4 * i + j] = m[4 * i] * o.m[j] + m[4 * i + 1] * o.m[4 + j] + m[4 * i + 2] * o.m[8 + j];
      r.m[4 * i + 3] = m[4 * i] * o.m[3] + m[4 * i + 1] * o.m[7] + m[4 * i + 2] * o.m[11] + m[4 * i + 3];
    }
    return r;
  }
  void apply(const double p[3], double out[3]) const {
    for (int i = 0; i < 3; ++i) out[i] = m[4 * i] * p[0] + m[4 * i + 1] * p[1] + m[4 * i + 2] * p[2] + m[4 * i + 3];
  }
};

class Context {
 public:
  Context(const hso_cam& cam, int device = 0, const hso_cfg* cfg = nullptr) : cam_(cam) {
    int rc = hso_create(device, &cam, cfg, &h_);
    if (rc != HSO_OK) throw std::runtime_error("hso_create failed (" + std::to_string(rc) + "): no sm_100 device? there is no CPU fallback");
  }
  // A context without a device: for host-only use of the data-model glue (makeDepthRef, getCloseViewObs, the wire formats) — e.g. unit tests on
  // a machine without a GPU. Every device call through it fails with HSO_ERR_INVALID (the C-ABI rejects a null context); nothing is computed on
  // the CPU instead.
  struct HostOnly {};
  Context(const hso_cam& cam, HostOnly) : cam_(cam) {}
  ~Context() { hso_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  hso_ctx* get() const { return h_; }
  const hso_cam& cam() const { return cam_; }
  void check(int rc) const {
    if (rc != HSO_OK) throw std::runtime_error(std::string("hso_b200: ") + hso_last_error(h_) + " (" + std::to_string(rc) + ")");
  }

 private:
  hso_ctx* h_ = nullptr;
  hso_cam cam_;
};

struct Frame;
struct Feature;

struct Point {  // include/hso/point.h:53-120 (subset the path reads)
  enum PointType { TYPE_DELETED = 0, TYPE_TEMPORARY = 1, TYPE_CANDIDATE = 2, TYPE_UNKNOWN = 3, TYPE_GOOD = 4 };
  enum FeatureType { FEATURE_GRADIENT = 0, FEATURE_EDGELET = 1, FEATURE_CORNER = 2 };
  double idist_ = 1.0;
  const Feature* hostFeature_ = nullptr;
  int type_ = TYPE_GOOD;
  int ftr_type_ = FEATURE_CORNER;
  double pos_[3] = {0, 0, 0};              // world position
  std::vector<const Feature*> obs_;        // keyframe observations (std::list<Feature*> in the reference)
  int n_failed_reproj_ = 0, n_succeeded_reproj_ = 0;
  bool isBad_ = false;
  // src/point.cpp:116-136 — pointer chasing over the observations stays on the host
  inline bool getCloseViewObs(const double framepos[3], const Feature*& ftr) const;
};

struct Feature {  // include/hso/feature.h:36-64 (subset)
  enum FeatureType { CORNER = 0, EDGELET = 1, GRADIENT = 2 };
  int type = CORNER;
  Frame* frame = nullptr;
  double px[2] = {0, 0};
  double f[3] = {0, 0, 1};
  int level = 0;
  Point* point = nullptr;
  double grad[2] = {1, 0};
};

struct Frame {  // include/hso/frame.h (subset): the pyramid lives on the device behind `id`
  Frame(Context& ctx, const uint8_t* img, int W, int H, int stride, double timestamp) : ctx_(ctx), timestamp_(timestamp) {
    // Frame::initFrame throws std::runtime_error when the image does not match the camera (src/frame.cpp:85-86)
    int rc = hso_frame_upload(ctx.get(), img, W, H, stride, &id, &integralImage_, &gradMean_);
    if (rc == HSO_ERR_INVALID) throw std::runtime_error("Frame: provided image has not the same size as the camera model or image is not grayscale");
    ctx.check(rc);
  }
  // A frame without a device pyramid (host-only contexts): pose, features and statistics are the caller's to fill.
  Frame(Context& ctx, Context::HostOnly, double timestamp) : ctx_(ctx), timestamp_(timestamp) {}
  ~Frame() { if (id >= 0) hso_frame_release(ctx_.get(), id); }
  Frame(const Frame&) = delete;
  Frame& operator=(const Frame&) = delete;
  Context& ctx_;
  hso_frame_id id = -1;
  double timestamp_;
  SE3 T_f_w_;
  float integralImage_ = 0, gradMean_ = 0, m_exposure_time = 1.f;
  std::vector<Feature> fts_;
  double Cov_[36] = {0};
  float m_error_in_px = 0;
  int keyFrameId_ = 0;
  void pos(double out[3]) const {  // Frame::pos() = T_f_w_.inverse().translation()
    const SE3 Ti = T_f_w_.inverse();
    out[0] = Ti.m[3]; out[1] = Ti.m[7]; out[2] = Ti.m[11];
  }
};
typedef std::shared_ptr<Frame> FramePtr;

inline bool Point::getCloseViewObs(const double framepos[3], const Feature*& ftr) const {
  if (obs_.empty()) return false;
  double od[3] = {framepos[0] - pos_[0], framepos[1] - pos_[1], framepos[2] - pos_[2]};
  const double on = std::sqrt(od[0] * od[0] + od[1] * od[1] + od[2] * od[2]);
  for (double& v : od) v /= on;
  size_t best = 0;
  double min_cos_angle = 0;
  for (size_t i = 0; i < obs_.size(); ++i) {
    double fp[3];
    obs_[i]->frame->pos(fp);
    double d[3] = {fp[0] - pos_[0], fp[1] - pos_[1], fp[2] - pos_[2]};
    const double dn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double cos_angle = (od[0] * d[0] + od[1] * d[1] + od[2] * d[2]) / dn;
    if (cos_angle > min_cos_angle) { min_cos_angle = cos_angle; best = i; }
  }
  ftr = obs_[best];
  return !(min_cos_angle < 0.5);  // observations more than 60 degrees apart are useless
}

class CoarseTracker {  // include/hso/CoarseTracker.h:134-143
 public:
  CoarseTracker(Context& ctx, bool inverse_composition, int max_level, int min_level, int n_iter, bool verbose = false)
      : ctx_(ctx), verbose_(verbose) {
    prm_.inverse_comp = inverse_composition ? 1 : 0; prm_.max_level = max_level; prm_.min_level = min_level; prm_.n_iter = n_iter;
  }

  // src/CoarseTracker.cpp:210-240 — stays on the host: it chases Feature -> Point -> host Feature -> host Frame pointers.
  static void makeDepthRef(const Frame& ref, std::vector<double>& dist) {
    dist.assign(ref.fts_.size(), -1.0);
    for (size_t i = 0; i < ref.fts_.size(); ++i) {
      const Feature& ft = ref.fts_[i];
      if (ft.point == nullptr) continue;
      const Feature* host = ft.point->hostFeature_;
      const double inv = 1.0 / ft.point->idist_;
      const double p_host[3] = {host->f[0] * inv, host->f[1] * inv, host->f[2] * inv};
      const SE3 T_r_h = ref.T_f_w_ * host->frame->T_f_w_.inverse();
      double p_ref[3];
      T_r_h.apply(p_host, p_ref);
      if (p_ref[2] < 0.00001) continue;
      dist[i] = std::sqrt(p_ref[0] * p_ref[0] + p_ref[1] * p_ref[1] + p_ref[2] * p_ref[2]);
    }
  }

  size_t run(FramePtr ref_frame, FramePtr cur_frame) {
    if (ref_frame->fts_.empty()) return 0;  // :53
    const size_t F = ref_frame->fts_.size();
    // the gather loop over the Feature list writes the compact layout of hso_track_job (32 B per feature): xyz = f * dist like :292, px as
    // float32 (exact for the tracker, see hso_b200.h), features without a usable depth left out like the reference skips them (:290)
    std::vector<double> xyz, dist;
    std::vector<float> px32;
    xyz.reserve(3 * F); px32.reserve(2 * F);
    makeDepthRef(*ref_frame, dist);
    for (size_t i = 0; i < F; ++i) {
      const Feature& ft = ref_frame->fts_[i];
      const double d = dist[i];
      if (!(d >= 0)) continue;
      px32.push_back((float)ft.px[0]); px32.push_back((float)ft.px[1]);
      xyz.push_back(ft.f[0] * d); xyz.push_back(ft.f[1] * d); xyz.push_back(ft.f[2] * d);
    }
    hso_track_job job;
    std::memset(&job, 0, sizeof job);
    job.ref = ref_frame->id; job.cur = cur_frame->id; job.n_features = (int32_t)(px32.size() / 2);
    job.xyz = xyz.data(); job.px32 = px32.data();
    const SE3 T_cur_ref = cur_frame->T_f_w_ * ref_frame->T_f_w_.inverse();  // :63
    std::memcpy(job.T_cur_ref, T_cur_ref.m, sizeof job.T_cur_ref);
    job.exposure_rat = cur_frame->integralImage_ / ref_frame->integralImage_;  // :60
    hso_track_result res;
    ctx_.check(hso_coarse_track(ctx_.get(), &prm_, &job, &res, nullptr, 0, nullptr));
    SE3 T;
    std::memcpy(T.m, res.T_cur_ref, sizeof T.m);
    cur_frame->T_f_w_ = T * ref_frame->T_f_w_;  // :198
    cur_frame->m_exposure_time = res.exposure_rat * ref_frame->m_exposure_time;
    if (res.exposure_rat > 0.99f && res.exposure_rat < 1.01f) cur_frame->m_exposure_time = ref_frame->m_exposure_time;  // :202
    last_ = res;
    return (size_t)res.n_tracked;
  }
  const hso_track_result& last_result() const { return last_; }

 private:
  Context& ctx_;
  hso_track_params prm_;
  bool verbose_;
  hso_track_result last_;
};

namespace pose_optimizer {
// include/hso/pose_optimizer.h:61-64. Outliers get Feature::point = NULL like src/pose_optimizer.cpp:721,735.
inline void optimizeLevenbergMarquardt3rd(Context& ctx, const double reproj_thresh, const size_t n_iter, const bool /*verbose*/, FramePtr& frame,
                                          double& estimated_scale, double& error_init, double& error_final, size_t& num_obs) {
  std::vector<double> f, p_host, grad, T_host;
  std::vector<int32_t> host_idx;
  std::vector<int8_t> level, ftype, ptype;
  std::vector<const Frame*> hosts;
  std::vector<size_t> idx;
  for (size_t i = 0; i < frame->fts_.size(); ++i) {
    const Feature& ft = frame->fts_[i];
    if (ft.point == nullptr) continue;
    const Feature* hf = ft.point->hostFeature_;
    const double inv = 1.0 / ft.point->idist_;
    for (int k = 0; k < 3; ++k) { f.push_back(ft.f[k]); p_host.push_back(hf->f[k] * inv); }
    grad.push_back(ft.grad[0]); grad.push_back(ft.grad[1]);
    size_t h = 0;
    for (; h < hosts.size(); ++h) if (hosts[h] == hf->frame) break;
    if (h == hosts.size()) { hosts.push_back(hf->frame); T_host.insert(T_host.end(), hf->frame->T_f_w_.m, hf->frame->T_f_w_.m + 12); }
    host_idx.push_back((int32_t)h);
    level.push_back((int8_t)ft.level); ftype.push_back((int8_t)ft.type); ptype.push_back((int8_t)ft.point->type_);
    idx.push_back(i);
  }
  std::vector<uint8_t> outlier(idx.size() + 1, 0);
  hso_pose_result res;
  ctx.check(hso_pose_optimize(ctx.get(), reproj_thresh, (int)n_iter, (int)frame->fts_.size(), (int)idx.size(), f.data(), p_host.data(), host_idx.data(),
                              (int)hosts.size(), T_host.data(), grad.data(), level.data(), ftype.data(), ptype.data(), frame->T_f_w_.m, outlier.data(),
                              &res));
  if (res.early_return) return;  // :456 leaves every output untouched
  std::memcpy(frame->T_f_w_.m, res.T_f_w, sizeof res.T_f_w);
  std::memcpy(frame->Cov_, res.cov, sizeof res.cov);
  for (size_t k = 0; k < idx.size(); ++k) if (outlier[k]) frame->fts_[idx[k]].point = nullptr;
  estimated_scale = res.estimated_scale; error_init = res.error_init; error_final = res.error_final; num_obs = (size_t)res.num_obs;
  frame->m_error_in_px = res.error_in_px;
}
}  // namespace pose_optimizer

// Detector part of FeatureExtractor::fastDetectST (src/feature_detection.cpp:498-523) on a device-resident level. The caller keeps
// haveFeatures_/getCellIndex bookkeeping and the octree distribution (host), exactly as after the reference's nonmax loop (:511-523).
struct KeyPointCandidate { int x, y; float shi_tomasi; int level; int fast_score; };
inline void fastDetectST(Context& ctx, const Frame& frame, int Level, float minThresh, std::vector<KeyPointCandidate>& out) {
  const int border = 8, scale = 1 << Level;
  const short fastThresh = (short)std::floor(minThresh);  // :502
  std::vector<hso_corner> buf(8192);
  int n = 0;
  ctx.check(hso_fast_detect(ctx.get(), frame.id, Level, fastThresh, border, buf.data(), (int)buf.size(), &n));
  if (n > (int)buf.size()) { buf.resize(n); ctx.check(hso_fast_detect(ctx.get(), frame.id, Level, fastThresh, border, buf.data(), n, &n)); }
  for (int i = 0; i < n; ++i) out.push_back(KeyPointCandidate{buf[i].x * scale, buf[i].y * scale, buf[i].shi_tomasi, Level, buf[i].score});  // :520-521
}

class Matcher {  // include/hso/matcher.h:113-153 (direct part)
 public:
  struct Options { int align_max_iter = 10; } options_;
  struct Candidate {  // what findMatchDirect knows after getCloseViewObs + getWarpMatrixAffine + getBestSearchLevel (matcher.cpp:276-309)
    const Feature* ref_ftr;
    double A_cur_ref[4];
    int search_level;
    double px_cur[2];  // in: initial estimate, out: refined
    bool found = false;
    double h_inv = 0;
  };
  explicit Matcher(Context& ctx) : ctx_(ctx) {}
  // src/matcher.cpp:74-85
  static int getBestSearchLevel(const double A[4], int max_level) {
    int search_level = 0;
    double D = A[0] * A[3] - A[1] * A[2];
    while (D > 3.0 && search_level < max_level) { search_level += 1; D *= 0.25; }
    return search_level;
  }
  // All candidates of one reprojectMap pass in one launch (the reference calls findMatchDirect once per candidate).
  void findMatchDirectBatch(std::vector<Candidate>& cands, Frame& cur_frame, int cur_keyFrameId = 0, const std::vector<int>* ref_keyFrameId = nullptr) {
    const size_t M = cands.size();
    if (M == 0) return;
    std::vector<hso_align_job> jobs(M);
    std::vector<hso_frame_id> refs(M);
    std::vector<hso_align_result> out(M);
    for (size_t m = 0; m < M; ++m) {
      const Feature* r = cands[m].ref_ftr;
      hso_align_job& j = jobs[m];
      std::memset(&j, 0, sizeof j);
      j.ref_level = r->level; j.search_level = cands[m].search_level; j.type = r->type;
      j.px_ref[0] = r->px[0]; j.px_ref[1] = r->px[1];
      std::memcpy(j.A_cur_ref, cands[m].A_cur_ref, sizeof j.A_cur_ref);
      j.grad[0] = r->grad[0]; j.grad[1] = r->grad[1];
      j.px_cur[0] = cands[m].px_cur[0]; j.px_cur[1] = cands[m].px_cur[1];
      // exposure scaling of the warped patch (src/matcher.cpp:317-330), LIGHT_THRESHOLD = 30
      const float a = cur_frame.m_exposure_time / r->frame->m_exposure_time;
      const bool near_kf = ref_keyFrameId ? (cur_keyFrameId - (*ref_keyFrameId)[m] < 4) : true;
      j.exposure_rat = a;
      j.scale_patch = (near_kf && std::fabs(a * 128 - 128) > 30.f) ? 1 : 0;
      refs[m] = r->frame->id;
    }
    ctx_.check(hso_align_batch(ctx_.get(), cur_frame.id, (int)M, jobs.data(), refs.data(), options_.align_max_iter, out.data()));
    for (size_t m = 0; m < M; ++m) {
      cands[m].found = out[m].ok != 0;
      cands[m].px_cur[0] = out[m].px_cur[0]; cands[m].px_cur[1] = out[m].px_cur[1];  // written even on failure (:372)
      cands[m].h_inv = out[m].h_inv;
    }
  }

 private:
  Context& ctx_;
};

// include/hso/reprojector.h — the data path of Reprojector::reprojectMap (src/reprojector.cpp:88-331) in ONE device call. The caller
// enumerates the points exactly like the reference's loops over the covisible / close keyframes, the candidates and the temporary points
// (those loops chase Map / Frame lists and stay on the host) and hands them over in that order.
class Reprojector {
 public:
  struct Grid { int cell_size = 0, grid_n_cols = 0, grid_n_rows = 0; std::vector<int32_t> cell_order; } grid_;
  size_t n_matches_ = 0, n_trials_ = 0;
  int nFeatures_ = 0;
  // Points the reference hands to map_.safeDeletePoint / point_candidates_.deleteCandidatePoint after a failed match (:369-372,565-568). The Map
  // is the caller's, so the deletions are surfaced here in the order the reference would issue them within a call.
  std::vector<Point*> points_to_delete_, candidates_to_delete_;
  Reprojector(Context& ctx, size_t max_fts) : ctx_(ctx), max_fts_(max_fts) {
    // Reprojector::initializeGrid (:58-77), caculateGridSize (:53-56)
    const int W = ctx.cam().width, H = ctx.cam().height;
    grid_.cell_size = (int)std::floor(std::sqrt((float)(W * H) / (float)max_fts) * 0.6);
    grid_.grid_n_cols = (int)std::ceil((double)W / grid_.cell_size);
    grid_.grid_n_rows = (int)std::ceil((double)H / grid_.cell_size);
    grid_.cell_order.resize((size_t)grid_.grid_n_cols * grid_.grid_n_rows);
    for (size_t i = 0; i < grid_.cell_order.size(); ++i) grid_.cell_order[i] = (int32_t)i;
  }
  // Reprojector::resetGrid (:79-86) re-shuffles grid_.cell_order with std::random_shuffle; the permutation is the caller's to choose
  // (pass any UniformRandomBitGenerator) so that runs are reproducible.
  template <class Rng>
  void resetGrid(Rng& rng) {
    n_matches_ = 0; n_trials_ = 0; nFeatures_ = 0;
    for (size_t i = grid_.cell_order.size(); i > 1; --i) std::swap(grid_.cell_order[i - 1], grid_.cell_order[rng() % i]);
  }
  // points: in the order the reference calls reprojectPoint. Side effects of reprojectCell / reprojectCellAll (:351-424,545-615) are applied
  // here: n_failed_reproj_ / n_succeeded_reproj_, TYPE_UNKNOWN -> TYPE_GOOD promotion, isBad_, and the new Features of `frame`.
  void reprojectMap(FramePtr frame, const std::vector<Point*>& points, std::vector<Frame*>& keyframes) {
    const size_t M = points.size();
    std::vector<hso_reproj_cand> cands(M);
    std::vector<const Feature*> ref_ftrs(M, nullptr);
    std::vector<double> T_f_w;
    auto pose_index = [&](const Frame* f) {
      for (size_t k = 0; k < keyframes.size(); ++k) if (keyframes[k] == f) return (int32_t)k;
      keyframes.push_back(const_cast<Frame*>(f));
      return (int32_t)(keyframes.size() - 1);
    };
    double cur_pos[3];
    frame->pos(cur_pos);
    for (size_t i = 0; i < M; ++i) {
      const Point* pt = points[i];
      hso_reproj_cand& c = cands[i];
      std::memset(&c, 0, sizeof c);
      const Feature* host = pt->hostFeature_;
      const double inv = 1.0 / pt->idist_;
      for (int k = 0; k < 3; ++k) c.p_host[k] = host->f[k] * inv;                       // :508
      c.host_pose = pose_index(host->frame);
      c.pt_type = pt->type_; c.pt_ftr_type = pt->ftr_type_;
      const Feature* ref = nullptr;
      c.ref_pose = -1;
      if (pt->getCloseViewObs(cur_pos, ref)) {                                            // matcher.cpp:276
        ref_ftrs[i] = ref;
        c.ref_pose = pose_index(ref->frame);
        c.ref_frame = ref->frame->id;
        c.ref_level = ref->level; c.ftr_type = ref->type;
        c.px_ref[0] = ref->px[0]; c.px_ref[1] = ref->px[1];
        for (int k = 0; k < 3; ++k) c.f_ref[k] = ref->f[k];
        c.grad[0] = ref->grad[0]; c.grad[1] = ref->grad[1];
        if (ref->frame == host->frame) {                                                  // matcher.cpp:298-309
          c.depth_ref = inv;
        } else {
          double rp[3];
          ref->frame->pos(rp);
          c.depth_ref = std::sqrt((rp[0] - pt->pos_[0]) * (rp[0] - pt->pos_[0]) + (rp[1] - pt->pos_[1]) * (rp[1] - pt->pos_[1]) +
                                  (rp[2] - pt->pos_[2]) * (rp[2] - pt->pos_[2]));
        }
        const float a = frame->m_exposure_time / ref->frame->m_exposure_time;           // matcher.cpp:317-321
        c.exposure_rat = a;
        c.scale_patch = (frame->keyFrameId_ - ref->frame->keyFrameId_ < 4 && std::fabs(a * 128 - 128) > 30.f) ? 1 : 0;
      }
    }
    for (Frame* kf : keyframes) T_f_w.insert(T_f_w.end(), kf->T_f_w_.m, kf->T_f_w_.m + 12);
    hso_reproj_grid g;
    g.cell_size = grid_.cell_size; g.n_cols = grid_.grid_n_cols; g.n_rows = grid_.grid_n_rows; g.max_fts = (int32_t)max_fts_;
    g.align_max_iter = 10; g.pad_ = 0;
    std::vector<hso_reproj_result> res(M + 1);
    hso_reproj_summary summ;
    ctx_.check(hso_reproject_match(ctx_.get(), frame->id, frame->T_f_w_.m, (int)keyframes.size(), T_f_w.data(), (int)M, cands.data(), &g,
                                   grid_.cell_order.data(), res.data(), &summ));
    n_matches_ = (size_t)summ.n_matches; n_trials_ = (size_t)summ.n_trials; nFeatures_ = summ.n_in_frame;
    points_to_delete_.clear(); candidates_to_delete_.clear();
    // new Features in the reference's order of creation
    std::vector<int> by_order(summ.n_matches, -1);
    for (size_t i = 0; i < M; ++i) {
      Point* pt = points[i];
      const hso_reproj_result& r = res[i];
      if (!r.tried) continue;
      if (!r.matched) {
        pt->n_failed_reproj_++;                                                           // :368-378
        if (pt->type_ == Point::TYPE_UNKNOWN && pt->n_failed_reproj_ > 15) points_to_delete_.push_back(pt);        // map_.safeDeletePoint
        if (pt->type_ == Point::TYPE_CANDIDATE && pt->n_failed_reproj_ > 30) candidates_to_delete_.push_back(pt);  // deleteCandidatePoint
        if (pt->type_ == Point::TYPE_TEMPORARY && pt->n_failed_reproj_ > 30) pt->isBad_ = true;
        continue;
      }
      pt->n_succeeded_reproj_++;
      if (pt->type_ == Point::TYPE_UNKNOWN && pt->n_succeeded_reproj_ > 10) pt->type_ = Point::TYPE_GOOD;  // :385-386
      if (r.order < 0 || r.order >= summ.n_matches || by_order[r.order] >= 0)
        throw std::runtime_error("hso_reproject_match: inconsistent creation order");  // would drop or overwrite a Feature
      by_order[r.order] = (int)i;
    }
    for (int i : by_order) {
      if (i < 0) continue;
      const hso_reproj_result& r = res[i];
      Feature nf;                                                                         // new Feature(frame.get(), it->px, matcher_.search_level_)
      nf.frame = frame.get(); nf.px[0] = r.px[0]; nf.px[1] = r.px[1]; nf.level = r.search_level; nf.point = points[i];
      const Feature* ref = ref_ftrs[i];
      if (ref->type == Feature::EDGELET) {                                                // :398-409
        nf.type = Feature::EDGELET;
        const double gx = r.A_cur_ref[0] * ref->grad[0] + r.A_cur_ref[1] * ref->grad[1], gy = r.A_cur_ref[2] * ref->grad[0] + r.A_cur_ref[3] * ref->grad[1];
        const double n = std::sqrt(gx * gx + gy * gy);
        nf.grad[0] = gx / n; nf.grad[1] = gy / n;
      } else {
        nf.type = ref->type == Feature::GRADIENT ? Feature::GRADIENT : Feature::CORNER;
      }
      frame->fts_.push_back(nf);
    }
  }

  // The seed stage of reprojectMap (src/reprojector.cpp:309-328): call after reprojectMap when n_matches_ < 100 and
  // Options::reproject_unconverged_seeds. `seeds`: the depth filter's seeds that pass the reference's own filter
  // (sqrt(sigma2) < z_range / reproject_seed_thresh && !haveReprojected, :316), in list order. For every accepted seed a TYPE_TEMPORARY Point and
  // a Feature are created exactly like Reprojector::reprojectorSeeds does (:438-489); the caller marks the seed haveReprojected / temp and hands
  // the point to map_.point_candidates_.addPauseSeedPoint (the Map is the caller's). Returns the accepted seeds' indices in creation order.
  struct SeedRef { const Feature* ftr; float mu, sigma2; };
  std::vector<int> reprojectSeeds(FramePtr frame, const std::vector<SeedRef>& seeds, std::vector<Frame*>& keyframes, std::vector<std::unique_ptr<Point>>& new_points) {
    const size_t S = seeds.size();
    std::vector<hso_seed_obs> rec(S);
    std::vector<double> T_f_w;
    auto pose_index = [&](const Frame* f) {
      for (size_t k = 0; k < keyframes.size(); ++k) if (keyframes[k] == f) return (int32_t)k;
      keyframes.push_back(const_cast<Frame*>(f));
      return (int32_t)(keyframes.size() - 1);
    };
    for (size_t i = 0; i < S; ++i) {
      const Feature* ft = seeds[i].ftr;
      hso_seed_obs& r = rec[i];
      std::memset(&r, 0, sizeof r);
      r.px[0] = ft->px[0]; r.px[1] = ft->px[1];
      for (int k = 0; k < 3; ++k) r.f[k] = ft->f[k];
      r.grad[0] = ft->grad[0]; r.grad[1] = ft->grad[1];
      r.ref_frame = ft->frame->id; r.ref_pose = pose_index(ft->frame);
      r.level = ft->level; r.ftr_type = ft->type;
      r.mu = seeds[i].mu; r.sigma2 = seeds[i].sigma2;
      r.exposure_rat = frame->m_exposure_time / ft->frame->m_exposure_time;                 // matcher.cpp:472
    }
    for (Frame* kf : keyframes) T_f_w.insert(T_f_w.end(), kf->T_f_w_.m, kf->T_f_w_.m + 12);
    hso_reproj_grid g;
    g.cell_size = grid_.cell_size; g.n_cols = grid_.grid_n_cols; g.n_rows = grid_.grid_n_rows; g.max_fts = (int32_t)max_fts_;
    g.align_max_iter = 10; g.pad_ = 0;
    std::vector<hso_reproj_result> res(S + 1);
    hso_reproj_summary summ;
    ctx_.check(hso_reproject_seeds(ctx_.get(), frame->id, frame->T_f_w_.m, (int)keyframes.size(), T_f_w.data(), (int)S, rec.data(), &g,
                                   grid_.cell_order.data(), (int)n_matches_, res.data(), &summ));
    const int n_new = summ.n_matches - (int)n_matches_;
    n_matches_ = (size_t)summ.n_matches;
    std::vector<int> by_order(n_new > 0 ? n_new : 0, -1);
    for (size_t i = 0; i < S; ++i)
      if (res[i].matched) {
        if (res[i].order < 0 || res[i].order >= n_new || by_order[res[i].order] >= 0) throw std::runtime_error("hso_reproject_seeds: inconsistent creation order");
        by_order[res[i].order] = (int)i;
      }
    for (int i : by_order) {
      if (i < 0) continue;
      const Feature* ft = seeds[i].ftr;
      const hso_reproj_result& r = res[i];
      // Point(xyz_world, seed.ftr): xyz_world = T_f_w^-1 * (f / mu); idist_ = mu; hostFeature_ = seed.ftr; TYPE_TEMPORARY (:444-452)
      std::unique_ptr<Point> pt(new Point());
      const double inv = 1.0 / seeds[i].mu;
      const double ph[3] = {ft->f[0] * inv, ft->f[1] * inv, ft->f[2] * inv};
      ft->frame->T_f_w_.inverse().apply(ph, pt->pos_);
      pt->idist_ = seeds[i].mu; pt->hostFeature_ = ft; pt->type_ = Point::TYPE_TEMPORARY;
      pt->ftr_type_ = ft->type == Feature::EDGELET ? Point::FEATURE_EDGELET : (ft->type == Feature::CORNER ? Point::FEATURE_CORNER : Point::FEATURE_GRADIENT);
      Feature nf;
      nf.frame = frame.get(); nf.px[0] = r.px[0]; nf.px[1] = r.px[1]; nf.level = r.search_level; nf.point = pt.get();
      if (ft->type == Feature::EDGELET) {
        nf.type = Feature::EDGELET;
        const double gx = r.A_cur_ref[0] * ft->grad[0] + r.A_cur_ref[1] * ft->grad[1], gy = r.A_cur_ref[2] * ft->grad[0] + r.A_cur_ref[3] * ft->grad[1];
        const double n = std::sqrt(gx * gx + gy * gy);
        nf.grad[0] = gx / n; nf.grad[1] = gy / n;
      } else {
        nf.type = ft->type == Feature::GRADIENT ? Feature::GRADIENT : Feature::CORNER;
      }
      frame->fts_.push_back(nf);
      new_points.push_back(std::move(pt));
    }
    return by_order;
  }

 private:
  Context& ctx_;
  size_t max_fts_;
};

// The front end of FrameHandlerMono::addImage for B independent streams in one chunk-pipelined call: new Frame(cam, img) for every image and
// CoarseTracker::run(ref_frames[b], new frame) (src/frame_handler_mono.cpp:92,190-204).
inline void addImagesAndTrack(Context& ctx, const std::vector<const uint8_t*>& imgs, int W, int H, int stride, const std::vector<FramePtr>& ref_frames,
                              const std::vector<SE3>& T_cur_ref_init, bool inverse_composition, int max_level, int min_level, int n_iter,
                              std::vector<hso_frame_id>& new_ids, std::vector<hso_track_result>& results) {
  const size_t B = imgs.size();
  std::vector<hso_track_job> jobs(B);
  // the gather loop over the Feature list writes the compact layout directly (32 B per feature: the call is bound by the host->device copies):
  // xyz = f * dist like src/CoarseTracker.cpp:292, px as float32 (exact, see hso_b200.h), features without a usable depth left out
  std::vector<std::vector<double>> xyz(B), dist(B);
  std::vector<std::vector<float>> px32(B);
  for (size_t b = 0; b < B; ++b) {
    const Frame& ref = *ref_frames[b];
    const size_t F = ref.fts_.size();
    xyz[b].reserve(3 * F); px32[b].reserve(2 * F);
    CoarseTracker::makeDepthRef(ref, dist[b]);
    for (size_t i = 0; i < F; ++i) {
      const double d = dist[b][i];
      if (!(d >= 0)) continue;
      px32[b].push_back((float)ref.fts_[i].px[0]); px32[b].push_back((float)ref.fts_[i].px[1]);
      for (int k = 0; k < 3; ++k) xyz[b].push_back(ref.fts_[i].f[k] * d);
    }
    hso_track_job& j = jobs[b];
    std::memset(&j, 0, sizeof j);
    j.ref = ref.id; j.n_features = (int32_t)(px32[b].size() / 2);
    j.xyz = xyz[b].data(); j.px32 = px32[b].data();
    std::memcpy(j.T_cur_ref, T_cur_ref_init[b].m, sizeof j.T_cur_ref);
    j.exposure_rat = -1.f;  // formed on the device from the two frames' integralImage_ (CoarseTracker.cpp:60)
  }
  hso_track_params prm;
  prm.inverse_comp = inverse_composition ? 1 : 0; prm.max_level = max_level; prm.min_level = min_level; prm.n_iter = n_iter;
  new_ids.resize(B); results.resize(B);
  ctx.check(hso_add_frames_track_batch(ctx.get(), &prm, (int)B, imgs.data(), W, H, stride, jobs.data(), new_ids.data(), nullptr, nullptr, results.data()));
}

// ---- wire / disk formats of the driver around the path (row N4) — plain host code ------------------------------------------------------
// ImageReader's timestamp file (src/ImageReader.cpp:30-66): four line formats tried in this order —
//   "stamp x y z a b c d" (TUM ground-truth style), "id stamp exposure" (TUM monoVO times.txt), "id stamp", "stamp".
// Returns false for a line none of them matches (the reference skips it). Quirk kept: the cascade is tried in that order, so a bare numeric
// stamp such as "1403636580.263555" is consumed by "%d %s" (id = 1403636580, stamp = ".263555") before the single-token format is reached.
inline bool parseTimestampLine(const char* buf, std::string& stamp_out) {
  int id;
  char stamp[100];
  float x, y, z, a, b, c, d, exposure = 0;
  if (8 == std::sscanf(buf, "%99s %f %f %f %f %f %f %f", stamp, &x, &y, &z, &a, &b, &c, &d)) { stamp_out = stamp; return true; }
  if (3 == std::sscanf(buf, "%d %99s %f", &id, stamp, &exposure)) { stamp_out = stamp; return true; }
  if (2 == std::sscanf(buf, "%d %99s", &id, stamp)) { stamp_out = stamp; return true; }
  if (1 == std::sscanf(buf, "%99s", stamp)) { stamp_out = stamp; return true; }
  return false;
}

// One line of BenchmarkNode::saveResult (test/test_dataset.cpp:312-335): "stamp tx ty tz qx qy qz qw" of T_w_f = T_f_w^-1 (TUM trajectory
// format; the keyframe id replaces the stamp when the sequence has no timestamps), default ostream precision like the reference.
inline void writeTrajectoryLine(std::ostream& os, bool stamp_valid, int id, const std::string& timestamp_s, const SE3& T_f_w) {
  const SE3 Tinv = T_f_w.inverse();
  const double* m = Tinv.m;
  // unit quaternion of the rotation block (Eigen's Quaternion(Matrix3) branches)
  double qw, qx, qy, qz;
  const double tr = m[0] + m[5] + m[10];
  if (tr > 0) {
    double r = std::sqrt(tr + 1.0);
    qw = 0.5 * r; r = 0.5 / r;
    qx = (m[9] - m[6]) * r; qy = (m[2] - m[8]) * r; qz = (m[4] - m[1]) * r;
  } else {
    const double mm[3][3] = {{m[0], m[1], m[2]}, {m[4], m[5], m[6]}, {m[8], m[9], m[10]}};
    int i = 0;
    if (mm[1][1] > mm[0][0]) i = 1;
    if (mm[2][2] > mm[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double r = std::sqrt(mm[i][i] - mm[j][j] - mm[k][k] + 1.0);
    double v[3];
    v[i] = 0.5 * r; r = 0.5 / r;
    qw = (mm[k][j] - mm[j][k]) * r;
    v[j] = (mm[j][i] + mm[i][j]) * r;
    v[k] = (mm[k][i] + mm[i][k]) * r;
    qx = v[0]; qy = v[1]; qz = v[2];
  }
  if (!stamp_valid) os << id << " "; else os << timestamp_s << " ";
  os << m[3] << " " << m[7] << " " << m[11] << " " << qx << " " << qy << " " << qz << " " << qw << std::endl;
}

}  // namespace b200
}  // namespace hso
