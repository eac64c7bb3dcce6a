// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or executed by the product path (hso_b200/).
//
// C entry points over the REFERENCE'S OWN CODE. oracle/Makefile compiles the reference's hot-path translation units unmodified, from where they
// lie under /root/reference (src/CoarseTracker.cpp, feature_alignment.cpp, matcher.cpp, pose_optimizer.cpp, frame.cpp, point.cpp, camera.cpp,
// config.cpp, src/vikit/{vision,robust_cost,math_utils}.cpp, thirdparty/Sophus/sophus/{so3,se3}.cpp), against the stand-in headers in
// oracle/shim/ for the three system libraries this image lacks (Eigen, OpenCV, Boost), into oracle/_ref/libhso_ref.so. This file is the only
// glue: it builds the reference's own objects (hso::Frame through its real constructor — pyramid, Sobel images, statistics —, hso::Feature,
// hso::Point, the camera models) from flat arrays and calls the reference's functions. Nothing here restates an algorithm; where a private
// member function is called directly (per-evaluation parity of the tracker) the file is compiled with -fno-access-control.
//
// Uses: tests/test_oracle_vs_reference.py pins the oracle restatement (oracle/*.cpp) to these functions; bench.py --impl reference and the
// cpu_baseline leg time them (kind = "reference").
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <tuple>
#include <vector>

#include <hso/CoarseTracker.h>
#include <hso/camera.h>
#include <hso/config.h>
#include <hso/depth_filter.h>
#include <hso/feature.h>
#include <hso/feature_alignment.h>
#include <hso/frame.h>
#include <hso/matcher.h>
#include <hso/point.h>
#include <hso/map.h>
#include <hso/pose_optimizer.h>
#include <hso/reprojector.h>
#include <hso/vikit/math_utils.h>
#include <hso/vikit/robust_cost.h>
#include <hso/vikit/vision.h>

#include "hso_oracle.h"

using namespace hso;

// The three symbols of src/feature_detection.cpp (not among the compiled TUs: cv::Canny, the octree distribution) that depth_filter.cpp references, so that
// the library links: the detector's grid bookkeeping, which observeDepthRow calls for keyframes only (:668-672), and the two calls of
// DepthFilter::initializeSeeds (:169-170). None is reached by ref_depth_observe — the active frames of the tests are not keyframes and no seeds are
// initialised through the filter — and each one aborts if it ever is.
void hso::feature_detection::FeatureExtractor::setGridOccpuancy(const Vector2d&, Feature*) { std::abort(); }
void hso::feature_detection::FeatureExtractor::setExistingFeatures(const Features&) { std::abort(); }
void hso::feature_detection::FeatureExtractor::detect(Frame*, const float, const float, Features&, Frame*) { std::abort(); }

namespace {

SE3 se3_from_rt(const double* rt) {
  Matrix3d R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R(i, j) = rt[4 * i + j];
  return SE3(R, Vector3d(rt[3], rt[7], rt[11]));
}
void se3_to_rt(const SE3& T, double* rt) {
  const Matrix3d R = T.rotation_matrix();
  const Vector3d t = T.translation();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) rt[4 * i + j] = R(i, j);
    rt[4 * i + 3] = t[i];
  }
}

// The reference's camera objects, built once per parameter set (PinholeCamera's constructor builds the undistortion maps).
AbstractCamera* get_cam(const orc_cam* c) {
  typedef std::tuple<int, int, int, int, double, double, double, double, double, double, double, double, double> Key;
  static std::map<Key, std::unique_ptr<AbstractCamera>> cache;
  const Key k(c->model, c->width, c->height, c->undistort, c->fx, c->fy, c->cx, c->cy, c->d[0], c->d[1], c->d[2], c->d[3], c->d[4]);
  auto it = cache.find(k);
  if (it != cache.end()) return it->second.get();
  AbstractCamera* cam = nullptr;
  if (c->model == 0) cam = new PinholeCamera(c->width, c->height, c->fx, c->fy, c->cx, c->cy, c->d[0], c->d[1], c->d[2], c->d[3], c->d[4]);
  else if (c->model == 1) cam = new FOVCamera(c->width, c->height, c->fx, c->fy, c->cx, c->cy, c->d[0], c->undistort != 0);
  else cam = new EquidistantCamera(c->width, c->height, c->fx, c->fy, c->cx, c->cy, c->d[0], c->d[1], c->d[2], c->d[3]);
  cache[k].reset(cam);
  return cam;
}

cv::Mat mat_from_u8(const uint8_t* img, int w, int h) {
  cv::Mat m(h, w, CV_8UC1);
  for (int y = 0; y < h; ++y) std::memcpy(m.ptr(y), img + (size_t)y * w, (size_t)w);
  return m;
}

// A reference Frame plus the Points this wrapper created for its features (the reference's Map owns points; here the handle does).
struct FrameHandle {
  FramePtr frame;
  std::vector<Point*> points;
  ~FrameHandle() { for (Point* p : points) delete p; }
};

void clear_features(FrameHandle* h) {
  for (Feature* f : h->frame->fts_) delete f;
  h->frame->fts_.clear();
  for (Point* p : h->points) delete p;
  h->points.clear();
}

// ref_frame->fts_ from flat arrays, every point hosted by its own feature in this frame: makeDepthRef (CoarseTracker.cpp:210-240) then yields
// |T_ref T_ref^-1 (f / idist)| = dist up to rounding, and the trackers' xyz_ref = f * dist is the oracle's.
void set_track_features(FrameHandle* h, int F, const double* px, const double* f, const double* dist) {
  clear_features(h);
  Frame* fr = h->frame.get();
  for (int i = 0; i < F; ++i) {
    Feature* ft = new Feature(fr, Vector2d(px[2 * i], px[2 * i + 1]), Vector3d(f[3 * i], f[3 * i + 1], f[3 * i + 2]), 0);
    if (dist[i] >= 0) {
      Point* pt = new Point(fr->T_f_w_.inverse() * (ft->f * dist[i]), ft);
      pt->hostFeature_ = ft;
      pt->idist_ = 1.0 / dist[i];
      pt->type_ = Point::TYPE_GOOD;
      pt->ftr_type_ = Point::FEATURE_CORNER;
      ft->point = pt;
      h->points.push_back(pt);
    }
    fr->fts_.push_back(ft);
  }
}

// what CoarseTracker::run does at the top of every level (src/CoarseTracker.cpp:76-92), for calling the per-level members directly
void enter_level(CoarseTracker& tr, int level) {
  tr.m_level = level;
  std::fill(tr.m_visible_fts.begin(), tr.m_visible_fts.end(), false);
  tr.m_offset_all = tr.m_max_level - tr.m_level + tr.m_pattern_offset;
  tr.HALF_PATCH_SIZE = tr.staticPatternPadding[tr.m_offset_all];
  tr.PATCH_AREA = tr.staticPatternNum[tr.m_offset_all];
  tr.m_ref_patch_cache = cv::Mat(tr.m_ref_frame->fts_.size(), tr.PATCH_AREA, CV_32F);
  tr.m_visible_fts.resize(tr.m_ref_frame->fts_.size(), false);
  tr.m_jacobian_cache_true.resize(Eigen::NoChange, tr.m_ref_patch_cache.rows * tr.PATCH_AREA);
  tr.m_jacobian_cache_raw.resize(Eigen::NoChange, tr.m_ref_patch_cache.rows * tr.PATCH_AREA);
  tr.precomputeReferencePatches();
}

}  // namespace

extern "C" {

const char* ref_describe() {
  return "luodongting/HSO reference sources compiled unmodified against oracle/shim (Eigen/OpenCV/Boost stand-ins): CoarseTracker, "
         "feature_alignment, matcher, pose_optimizer, frame, point, camera, config, vikit/{vision,robust_cost,math_utils}, Sophus so3/se3";
}

// ---- a11: camera models (src/camera.cpp) ------------------------------------------------------------------------------------------------
void ref_world2cam(const orc_cam* cam, const double xyz[3], double px_out[2]) {
  const Vector2d px = get_cam(cam)->world2cam(Vector3d(xyz[0], xyz[1], xyz[2]));
  px_out[0] = px[0]; px_out[1] = px[1];
}
void ref_cam2world(const orc_cam* cam, double u, double v, double xyz_out[3]) {
  const Vector3d f = get_cam(cam)->cam2world(u, v);
  xyz_out[0] = f[0]; xyz_out[1] = f[1]; xyz_out[2] = f[2];
}
double ref_error_multiplier2(const orc_cam* cam) { return get_cam(cam)->errorMultiplier2(); }

// ---- a12: Sophus (thirdparty/Sophus/sophus/se3.cpp, so3.cpp) -------------------------------------------------------------------------------
void ref_se3_exp(const double tangent[6], double rt_out[12]) {
  Matrix<double, 6, 1> t;
  for (int i = 0; i < 6; ++i) t[i] = tangent[i];
  se3_to_rt(SE3::exp(t), rt_out);
}
void ref_se3_log(const double rt[12], double tangent_out[6]) {
  const Matrix<double, 6, 1> t = se3_from_rt(rt).log();
  for (int i = 0; i < 6; ++i) tangent_out[i] = t[i];
}
void ref_se3_mul(const double a[12], const double b[12], double out[12]) { se3_to_rt(se3_from_rt(a) * se3_from_rt(b), out); }
void ref_se3_inverse(const double a[12], double out[12]) { se3_to_rt(se3_from_rt(a).inverse(), out); }
void ref_se3_apply(const double a[12], const double p[3], double out[3]) {
  const Vector3d q = se3_from_rt(a) * Vector3d(p[0], p[1], p[2]);
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2];
}

// ---- a17: robust cost, median (src/vikit/robust_cost.cpp:67-74,129-148; include/hso/vikit/math_utils.h:119-126) ----------------------------
float ref_mad_scale(const float* errors, int n) {
  std::vector<float> e(errors, errors + n);
  robust_cost::MADScaleEstimator est;
  return est.compute(e);
}
float ref_huber_weight(float x) {
  robust_cost::HuberWeightFunction w;
  return w.value(x);
}
float ref_get_median_f(const float* v, int n) {
  std::vector<float> e(v, v + n);
  return hso::getMedian(e);
}
double ref_get_median_d(const double* v, int n) {
  std::vector<double> e(v, v + n);
  return hso::getMedian(e);
}

// ---- a1: halfSample / createImgPyramid (src/vikit/vision.cpp:19-108, src/frame.cpp:296-314) ----------------------------------------------
void ref_half_sample(const uint8_t* in, int w, int h, uint8_t* out) {
  cv::Mat src = mat_from_u8(in, w, h), dst(h / 2, w / 2, CV_8U);
  hso::halfSample(src, dst);
  for (int y = 0; y < h / 2; ++y) std::memcpy(out + (size_t)y * (w / 2), dst.ptr(y), (size_t)(w / 2));
}
// out: levels concatenated tightly; lw / lh receive the level sizes. Returns the total number of bytes written.
int ref_create_pyramid(const uint8_t* img, int W, int H, int n_levels, uint8_t* out, int* lw, int* lh) {
  ImgPyr pyr;
  frame_utils::createImgPyramid(mat_from_u8(img, W, H), n_levels, pyr);
  size_t o = 0;
  for (int l = 0; l < n_levels; ++l) {
    lw[l] = pyr[l].cols; lh[l] = pyr[l].rows;
    for (int y = 0; y < pyr[l].rows; ++y) { std::memcpy(out + o, pyr[l].ptr(y), (size_t)pyr[l].cols); o += (size_t)pyr[l].cols; }
  }
  return (int)o;
}

// ---- a7: Accumulator7 (include/hso/MatrixAccumulator.h:29-141) exactly as computeGS drives it ----------------------------------------------
void ref_accumulator7(int n, const float* J /*7n*/, const float* w /*n*/, float* H49) {
  Accumulator7 acc;
  acc.initialize();
  for (int i = 0; i < n; ++i) acc.updateSingleWeighted(J[7 * i], J[7 * i + 1], J[7 * i + 2], J[7 * i + 3], J[7 * i + 4], J[7 * i + 5], J[7 * i + 6], w[i], 0);
  acc.finish();
  for (int r = 0; r < 7; ++r)
    for (int c = 0; c < 7; ++c) H49[7 * r + c] = acc.H(r, c);
}

// ---- a1 + a2: hso::Frame through its real constructor (src/frame.cpp:45-96,205-246) ---------------------------------------------------------
// Returns NULL when the constructor throws (wrong size / type, frame.cpp:85-86).
void* ref_frame_new(const orc_cam* cam, const uint8_t* img, int W, int H, const double T_f_w[12], double exposure_time, int keyframe_id) {
  try {
    std::unique_ptr<FrameHandle> h(new FrameHandle());
    h->frame.reset(new Frame(get_cam(cam), mat_from_u8(img, W, H), 0.0));
    if (T_f_w) h->frame->T_f_w_ = se3_from_rt(T_f_w);
    h->frame->m_exposure_time = exposure_time;
    h->frame->keyFrameId_ = keyframe_id;
    return h.release();
  } catch (const std::exception&) {
    return nullptr;
  }
}
void ref_frame_free(void* handle) {
  FrameHandle* h = (FrameHandle*)handle;
  if (!h) return;
  // ~Frame deletes its Features (frame.cpp:54-72); the Points are the handle's
  delete h;
}
void ref_frame_stats(void* handle, float* integral, float* grad_mean) {
  Frame* f = ((FrameHandle*)handle)->frame.get();
  *integral = f->integralImage_;
  *grad_mean = f->gradMean_;
}
int ref_frame_n_levels(void* handle) { return (int)((FrameHandle*)handle)->frame->img_pyr_.size(); }
void ref_frame_level_size(void* handle, int level, int* w, int* h) {
  const cv::Mat& m = ((FrameHandle*)handle)->frame->img_pyr_[level];
  *w = m.cols; *h = m.rows;
}
void ref_frame_level(void* handle, int level, uint8_t* out) {
  const cv::Mat& m = ((FrameHandle*)handle)->frame->img_pyr_[level];
  for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
}
void ref_frame_sobel(void* handle, int level, int16_t* gx, int16_t* gy) {
  Frame* f = ((FrameHandle*)handle)->frame.get();
  const cv::Mat &sx = f->sobelX_[level], &sy = f->sobelY_[level];
  for (int y = 0; y < sx.rows; ++y) {
    std::memcpy(gx + (size_t)y * sx.cols, sx.ptr(y), (size_t)sx.cols * 2);
    std::memcpy(gy + (size_t)y * sy.cols, sy.ptr(y), (size_t)sy.cols * 2);
  }
}
void ref_frame_set_pose(void* handle, const double T_f_w[12]) { ((FrameHandle*)handle)->frame->T_f_w_ = se3_from_rt(T_f_w); }
void ref_frame_get_pose(void* handle, double T_f_w[12]) { se3_to_rt(((FrameHandle*)handle)->frame->T_f_w_, T_f_w); }
void ref_frame_set_track_features(void* handle, int F, const double* px, const double* f, const double* dist) {
  set_track_features((FrameHandle*)handle, F, px, f, dist);
}

// ---- a3-a10: CoarseTracker::run (src/CoarseTracker.cpp:51-208) ----------------------------------------------------------------------------
// ref must carry features (ref_frame_set_track_features). The initial pose is T_cur_ref_io (written into cur.T_f_w_ relative to the reference
// frame's pose); the initial exposure ratio is what run() forms itself, cur.integralImage_ / ref.integralImage_ (:60). Returns run()'s value.
uint64_t ref_coarse_track(void* ref_handle, void* cur_handle, const orc_track_params* prm, double T_cur_ref_io[12], float* a_out,
                          double* exposure_time_out) {
  FrameHandle* r = (FrameHandle*)ref_handle;
  FrameHandle* c = (FrameHandle*)cur_handle;
  c->frame->T_f_w_ = se3_from_rt(T_cur_ref_io) * r->frame->T_f_w_;
  CoarseTracker tracker(prm->inverse_comp != 0, prm->max_level, prm->min_level, prm->n_iter, false);
  const size_t n = tracker.run(r->frame, c->frame);
  se3_to_rt(c->frame->T_f_w_ * r->frame->T_f_w_.inverse(), T_cur_ref_io);
  if (a_out) *a_out = r->frame->fts_.empty() ? 0.f : tracker.m_exposure_rat;
  if (exposure_time_out) *exposure_time_out = c->frame->m_exposure_time;
  return (uint64_t)n;
}

// One residual evaluation + normal equations at a given state with thresholds supplied: precomputeReferencePatches (:416-497),
// computeResiduals (:242-414) and computeGS (:499-525) called directly on a CoarseTracker set up like run() sets it up for `level`.
void ref_track_eval(void* ref_handle, void* cur_handle, int inverse_comp, int level, int max_level, const double T[12], float a, float huber,
                    float outlier, double H_out[49], double b_out[7], double* energy_out, int* total_terms, int* saturated_terms) {
  FrameHandle* r = (FrameHandle*)ref_handle;
  FrameHandle* c = (FrameHandle*)cur_handle;
  CoarseTracker tr(inverse_comp != 0, max_level, level, 0, false);
  tr.m_ref_frame = r->frame; tr.m_cur_frame = c->frame;
  tr.m_exposure_rat = a; tr.m_b = 0;
  tr.makeDepthRef();
  enter_level(tr, level);
  tr.m_huber_thresh = huber;
  tr.m_outlier_thresh = outlier;
  const double cutoff_error = tr.m_outlier_thresh;  // :100
  const double E = tr.computeResiduals(se3_from_rt(T), a, cutoff_error);
  Eigen::Matrix<double, 7, 7> H;
  Eigen::Matrix<double, 7, 1> b;
  tr.computeGS(H, b);
  for (int i = 0; i < 7; ++i) {
    for (int j = 0; j < 7; ++j) H_out[7 * i + j] = H(i, j);
    b_out[i] = b[i];
  }
  *energy_out = E;
  *total_terms = tr.m_total_terms;
  *saturated_terms = tr.m_saturated_terms;
}

// selectRobustFunctionLevel (:530-644) at a given state. n_errors is not observable in the reference (a local vector); -1 is returned there.
void ref_track_select_robust(void* ref_handle, void* cur_handle, int level, int max_level, const double T[12], float a, float* huber_out,
                             float* outlier_out) {
  FrameHandle* r = (FrameHandle*)ref_handle;
  FrameHandle* c = (FrameHandle*)cur_handle;
  CoarseTracker tr(false, max_level, level, 0, false);
  tr.m_ref_frame = r->frame; tr.m_cur_frame = c->frame;
  tr.makeDepthRef();
  enter_level(tr, level);
  tr.selectRobustFunctionLevel(se3_from_rt(T), a);
  *huber_out = tr.m_huber_thresh;
  *outlier_out = tr.m_outlier_thresh;
}

// makeDepthRef alone (:210-240) on a frame whose features' points are hosted in other frames: host_handles[i] is the frame hosting feature i's
// point (NULL: feature without point), f_host / idist the host bearing and inverse depth.
void ref_make_depth_ref(void* ref_handle, int F, void* const* host_handles, const double* f_host, const double* idist, double* dist_out) {
  FrameHandle* r = (FrameHandle*)ref_handle;
  clear_features(r);
  std::vector<std::unique_ptr<Feature>> host_fts;
  for (int i = 0; i < F; ++i) {
    Feature* ft = new Feature(r->frame.get(), Vector2d(0, 0), Vector3d(0, 0, 1), 0);
    if (host_handles[i]) {
      Frame* hf = ((FrameHandle*)host_handles[i])->frame.get();
      host_fts.emplace_back(new Feature(hf, Vector2d(0, 0), Vector3d(f_host[3 * i], f_host[3 * i + 1], f_host[3 * i + 2]), 0));
      Point* pt = new Point(Vector3d(0, 0, 0), host_fts.back().get());
      pt->hostFeature_ = host_fts.back().get();
      pt->idist_ = idist[i];
      ft->point = pt;
      r->points.push_back(pt);
    }
    r->frame->fts_.push_back(ft);
  }
  CoarseTracker tr(false, 4, 1, 0, false);
  tr.m_ref_frame = r->frame;
  tr.makeDepthRef();
  for (int i = 0; i < F; ++i) dist_out[i] = tr.m_pt_ref[i];
  clear_features(r);
}

// ---- a14, a15: feature_alignment float overloads (src/feature_alignment.cpp:164-308,464-605) -------------------------------------------------
int ref_align2d(const uint8_t* cur_img, int cols, int rows, int stride, const float* ref_patch_with_border, const float* ref_patch, int n_iter,
                double px_io[2], float* cur_patch_out) {
  const cv::Mat img(rows, cols, CV_8UC1, (void*)cur_img, (size_t)stride);
  Vector2d px(px_io[0], px_io[1]);
  float patch[64];
  const bool ok = feature_alignment::align2D(img, const_cast<float*>(ref_patch_with_border), const_cast<float*>(ref_patch), n_iter, px, false, patch);
  px_io[0] = px[0]; px_io[1] = px[1];
  if (cur_patch_out) std::memcpy(cur_patch_out, patch, sizeof patch);
  return ok ? 1 : 0;
}
int ref_align1d(const uint8_t* cur_img, int cols, int rows, int stride, const float dir[2], const float* ref_patch_with_border, const float* ref_patch,
                int n_iter, double px_io[2], double* h_inv_out, float* cur_patch_out) {
  const cv::Mat img(rows, cols, CV_8UC1, (void*)cur_img, (size_t)stride);
  Vector2d px(px_io[0], px_io[1]);
  float patch[64];
  double h_inv = 0;
  const bool ok = feature_alignment::align1D(img, Vector2f(dir[0], dir[1]), const_cast<float*>(ref_patch_with_border), const_cast<float*>(ref_patch), n_iter,
                                             px, h_inv, patch);
  px_io[0] = px[0]; px_io[1] = px[1];
  if (h_inv_out) *h_inv_out = h_inv;
  if (cur_patch_out) std::memcpy(cur_patch_out, patch, sizeof patch);
  return ok ? 1 : 0;
}

// ---- a13 pieces: warp (src/matcher.cpp:46-155) ----------------------------------------------------------------------------------------------
void ref_get_warp_matrix_affine(const orc_cam* cam, const double px_ref[2], const double f_ref[3], double depth_ref, const double T_cur_ref[12],
                                int level_ref, double A_cur_ref[4]) {
  Matrix2d A;
  AbstractCamera* c = get_cam(cam);
  warp::getWarpMatrixAffine(*c, *c, Vector2d(px_ref[0], px_ref[1]), Vector3d(f_ref[0], f_ref[1], f_ref[2]), depth_ref, se3_from_rt(T_cur_ref), level_ref, A);
  A_cur_ref[0] = A(0, 0); A_cur_ref[1] = A(0, 1); A_cur_ref[2] = A(1, 0); A_cur_ref[3] = A(1, 1);
}
int ref_get_best_search_level(const double A_cur_ref[4], int max_level) {
  Matrix2d A;
  A(0, 0) = A_cur_ref[0]; A(0, 1) = A_cur_ref[1]; A(1, 0) = A_cur_ref[2]; A(1, 1) = A_cur_ref[3];
  return warp::getBestSearchLevel(A, max_level);
}
void ref_warp_affine(const double A_cur_ref[4], const uint8_t* img_ref, int cols, int rows, const double px_ref[2], int level_ref, int search_level,
                     int halfpatch_size, float* patch_out) {
  Matrix2d A;
  A(0, 0) = A_cur_ref[0]; A(0, 1) = A_cur_ref[1]; A(1, 0) = A_cur_ref[2]; A(1, 1) = A_cur_ref[3];
  const cv::Mat img(rows, cols, CV_8UC1, (void*)img_ref);
  warp::warpAffine(A, img, Vector2d(px_ref[0], px_ref[1]), level_ref, search_level, halfpatch_size, patch_out);
}

// ---- a13 / a13b: whole Matcher::findMatchDirect (src/matcher.cpp:270-375) and Matcher::findMatchSeed (:442-518) for a list of candidates ------
// Candidates use the record of row N1 (orc_reproj_cand): the point is rebuilt as a reference Point with ONE observation — the feature
// (px_ref, f_ref, level, type, grad) in keyframe kf_handles[ref_frame] — hosted by a feature with bearing p_host / |p_host| and inverse depth
// 1 / |p_host| in keyframe kf_handles[host_pose]; its world position is T_host^-1 p_host. px_io: the reprojected pixel in, the pixel
// findMatchDirect leaves out. seed_mode != 0 runs findMatchSeed with Seed{ftr = that observation, mu = 1 / depth_ref} instead.
void ref_find_match_batch(void* cur_handle, int n_kf, void* const* kf_handles, int M, const orc_reproj_cand* cands, int seed_mode, double* px_io /*2M*/,
                          int32_t* ok_out, int32_t* search_level_out, double* A_out /*4M*/, double* h_inv_out /*M or NULL*/) {
  FrameHandle* cur = (FrameHandle*)cur_handle;
  Matcher matcher;
  for (int i = 0; i < M; ++i) {
    const orc_reproj_cand& c = cands[i];
    ok_out[i] = 0; search_level_out[i] = 0;
    for (int k = 0; k < 4; ++k) A_out[4 * i + k] = 0;
    if (c.ref_pose < 0 || c.ref_pose >= n_kf || c.host_pose < 0 || c.host_pose >= n_kf) continue;  // getCloseViewObs would have nothing to return
    Frame* kf_ref = ((FrameHandle*)kf_handles[c.ref_frame])->frame.get();
    Frame* kf_host = ((FrameHandle*)kf_handles[c.host_pose])->frame.get();
    const Vector3d p_host(c.p_host[0], c.p_host[1], c.p_host[2]);
    Feature obs(kf_ref, Vector2d(c.px_ref[0], c.px_ref[1]), Vector3d(c.f_ref[0], c.f_ref[1], c.f_ref[2]), c.ref_level);
    obs.type = (Feature::FeatureType)c.ftr_type;
    obs.grad = Vector2d(c.grad[0], c.grad[1]);
    Feature host(kf_host, Vector2d(0, 0), p_host * (1.0 / p_host.norm()), 0);
    Vector2d px(px_io[2 * i], px_io[2 * i + 1]);
    bool ok;
    if (!seed_mode) {
      Point pt(kf_host->T_f_w_.inverse() * p_host, &obs);
      pt.hostFeature_ = (c.ref_frame == c.host_pose) ? &obs : &host;
      pt.idist_ = (c.ref_frame == c.host_pose) ? 1.0 / c.depth_ref : 1.0 / p_host.norm();
      pt.type_ = (Point::PointType)c.pt_type;
      ok = matcher.findMatchDirect(pt, *cur->frame, px);
    } else {
      alignas(16) static unsigned char storage[sizeof(Seed)];  // findMatchSeed reads seed.ftr and seed.mu only; Seed's constructor lives in depth_filter.cpp
      std::memset(storage, 0, sizeof storage);
      Seed* seed = reinterpret_cast<Seed*>(storage);
      seed->ftr = &obs;
      seed->mu = (float)(1.0 / c.depth_ref);
      ok = matcher.findMatchSeed(*seed, *cur->frame, px);
    }
    px_io[2 * i] = px[0]; px_io[2 * i + 1] = px[1];
    ok_out[i] = ok ? 1 : 0;
    search_level_out[i] = matcher.search_level_;
    A_out[4 * i] = matcher.A_cur_ref_(0, 0); A_out[4 * i + 1] = matcher.A_cur_ref_(0, 1);
    A_out[4 * i + 2] = matcher.A_cur_ref_(1, 0); A_out[4 * i + 3] = matcher.A_cur_ref_(1, 1);
    if (h_inv_out) h_inv_out[i] = matcher.h_inv_;
  }
}

// a13b with the seed record of rows N3 / a13b (orc_seed_obs): Matcher::findMatchSeed(seed, frame, px) with Seed{ftr = the feature (px, f, level,
// type, grad) in keyframe kf_handles[ref_frame], mu}. px_io: the pixel Reprojector::reprojectorSeed computed in, the pixel findMatchSeed leaves out.
void ref_find_match_seed_batch(void* cur_handle, int n_kf, void* const* kf_handles, int S, const orc_seed_obs* seeds, double* px_io /*2S*/,
                               int32_t* ok_out, int32_t* search_level_out, double* A_out /*4S*/) {
  FrameHandle* cur = (FrameHandle*)cur_handle;
  Matcher matcher;
  alignas(16) static unsigned char storage[sizeof(Seed)];  // findMatchSeed reads seed.ftr and seed.mu only; Seed's constructor lives in depth_filter.cpp
  for (int i = 0; i < S; ++i) {
    const orc_seed_obs& s = seeds[i];
    ok_out[i] = 0; search_level_out[i] = 0;
    for (int k = 0; k < 4; ++k) A_out[4 * i + k] = 0;
    if (s.ref_frame < 0 || s.ref_frame >= n_kf) continue;
    Frame* kf = ((FrameHandle*)kf_handles[s.ref_frame])->frame.get();
    Feature obs(kf, Vector2d(s.px[0], s.px[1]), Vector3d(s.f[0], s.f[1], s.f[2]), s.level);
    obs.type = (Feature::FeatureType)s.ftr_type;
    obs.grad = Vector2d(s.grad[0], s.grad[1]);
    std::memset(storage, 0, sizeof storage);
    Seed* seed = reinterpret_cast<Seed*>(storage);
    seed->ftr = &obs;
    seed->mu = s.mu;
    seed->sigma2 = s.sigma2;
    matcher.A_cur_ref_.setZero();
    Vector2d px(px_io[2 * i], px_io[2 * i + 1]);
    const bool ok = matcher.findMatchSeed(*seed, *cur->frame, px);
    px_io[2 * i] = px[0]; px_io[2 * i + 1] = px[1];
    ok_out[i] = ok ? 1 : 0;
    search_level_out[i] = matcher.search_level_;
    A_out[4 * i] = matcher.A_cur_ref_(0, 0); A_out[4 * i + 1] = matcher.A_cur_ref_(0, 1);
    A_out[4 * i + 2] = matcher.A_cur_ref_(1, 0); A_out[4 * i + 3] = matcher.A_cur_ref_(1, 1);
  }
}

// ---- N1: the grid stage of Reprojector::reprojectMap through the reference's own member functions ---------------------------------------------
// Candidates use the record of row N1 (orc_reproj_cand), rebuilt as reference Points like in ref_find_match_batch (one observation = the chosen
// reference feature, so that Point::getCloseViewObs returns it). Per candidate, in list order, Reprojector::reprojectPoint (:504-529: projection,
// 8-px frame test, grid cell) — where reprojectMap's enumeration loops (:121-250, host side of the boundary) call it; then the selection, i.e.
// lines :260-303 restated verbatim below around the reference's reprojectCellAll (:545-615) / reprojectCell (:351-424: per-cell stable sort by
// point quality, first match wins, ++n_trials_ ahead of the TYPE_DELETED test, failure counters, feature creation). The grid is the reference's
// own (initializeGrid with Config::maxFts() = grid->max_fts; its geometry must equal *grid, else n_matches = -1), cell_order replaces the
// std::random_shuffle. Results: in_frame / cell / px per candidate, matched + creation order + search level from the Features the walk added to the
// frame, tried from the points' own success / failure counters.
void ref_reproject_match(void* cur_handle, int n_kf, void* const* kf_handles, int M, const orc_reproj_cand* cands, const orc_reproj_grid* grid,
                         const int32_t* cell_order, orc_reproj_result* out, orc_reproj_summary* summary) {
  FrameHandle* cur = (FrameHandle*)cur_handle;
  FramePtr frame = cur->frame;
  std::memset(summary, 0, sizeof *summary);
  for (int i = 0; i < M; ++i) { std::memset(&out[i], 0, sizeof out[i]); out[i].order = -1; out[i].cell = -1; }
  Config::maxFts() = (size_t)grid->max_fts;
  Map map;
  Reprojector rep(frame->cam_, map);
  if (rep.grid_.cell_size != grid->cell_size || rep.grid_.grid_n_cols != grid->n_cols || rep.grid_.grid_n_rows != grid->n_rows) { summary->n_matches = -1; return; }
  rep.resetGrid();
  rep.grid_.cell_order.assign(cell_order, cell_order + rep.grid_.cells.size());
  rep.matcher_.options_.align_max_iter = grid->align_max_iter;
  std::vector<std::unique_ptr<Feature>> feats;
  std::vector<std::unique_ptr<Point>> pts(M);
  std::map<Point*, int> index;
  std::vector<std::pair<Vector2d, Point*>> allPixelToDistribute;
  for (int i = 0; i < M; ++i) {
    const orc_reproj_cand& c = cands[i];
    if (c.host_pose < 0 || c.host_pose >= n_kf) continue;
    Frame* kf_host = ((FrameHandle*)kf_handles[c.host_pose])->frame.get();
    const Vector3d p_host(c.p_host[0], c.p_host[1], c.p_host[2]);
    feats.emplace_back(new Feature(kf_host, Vector2d(0, 0), p_host * (1.0 / p_host.norm()), 0));
    Feature* host = feats.back().get();
    Feature* obs = nullptr;
    if (c.ref_pose >= 0 && c.ref_pose < n_kf) {
      Frame* kf_ref = ((FrameHandle*)kf_handles[c.ref_frame])->frame.get();
      feats.emplace_back(new Feature(kf_ref, Vector2d(c.px_ref[0], c.px_ref[1]), Vector3d(c.f_ref[0], c.f_ref[1], c.f_ref[2]), c.ref_level));
      obs = feats.back().get();
      obs->type = (Feature::FeatureType)c.ftr_type;
      obs->grad = Vector2d(c.grad[0], c.grad[1]);
    }
    Point* pt = new Point(kf_host->T_f_w_.inverse() * p_host, obs ? obs : host);
    if (!obs) pt->obs_.clear();  // no observation in view: getCloseViewObs fails, findMatchDirect returns false (matcher.cpp:276)
    pt->hostFeature_ = (obs && c.ref_frame == c.host_pose) ? obs : host;
    pt->idist_ = (obs && c.ref_frame == c.host_pose) ? 1.0 / c.depth_ref : 1.0 / p_host.norm();
    pt->type_ = (Point::PointType)c.pt_type;
    pt->ftr_type_ = (Point::FeatureType)c.pt_ftr_type;
    pts[i].reset(pt);
    index[pt] = i;
    const size_t before = allPixelToDistribute.size();
    if (rep.reprojectPoint(frame, pt, allPixelToDistribute)) {
      const Vector2d& px = allPixelToDistribute[before].first;
      out[i].in_frame = 1;
      out[i].px[0] = px[0]; out[i].px[1] = px[1];
      out[i].cell = static_cast<int>(px[1] / rep.grid_.cell_size) * rep.grid_.grid_n_cols + static_cast<int>(px[0] / rep.grid_.cell_size);
    }
  }
  summary->n_in_frame = (int32_t)allPixelToDistribute.size();
  const size_t n_fts0 = frame->fts_.size();
  // ---- src/reprojector.cpp:260-303 ----
  if (allPixelToDistribute.size() < Config::maxFts() + 50) {
    summary->used_cell_all = 1;
    rep.reprojectCellAll(allPixelToDistribute, frame);
  } else {
    for (size_t i = 0; i < rep.grid_.cells.size(); ++i) {
      if (rep.reprojectCell(*rep.grid_.cells.at(rep.grid_.cell_order[i]), frame, false, false)) ++rep.n_matches_;
      if (rep.n_matches_ >= (size_t)Config::maxFts()) break;
    }
    if (rep.n_matches_ < (size_t)Config::maxFts()) {
      for (size_t i = rep.grid_.cells.size() - 1; i > 0; --i) {
        if (rep.reprojectCell(*rep.grid_.cells.at(rep.grid_.cell_order[i]), frame, true, false)) ++rep.n_matches_;
        if (rep.n_matches_ >= (size_t)Config::maxFts()) break;
      }
    }
    if (rep.n_matches_ < (size_t)Config::maxFts()) {
      for (size_t i = 0; i < rep.grid_.cells.size(); ++i) {
        rep.reprojectCell(*rep.grid_.cells.at(rep.grid_.cell_order[i]), frame, true, true);
        if (rep.n_matches_ >= (size_t)Config::maxFts()) break;
      }
    }
  }
  summary->n_matches = (int32_t)rep.n_matches_;
  summary->n_trials = (int32_t)rep.n_trials_;
  int order = 0;
  size_t k = 0;
  for (auto it = frame->fts_.begin(); it != frame->fts_.end(); ++it, ++k) {
    if (k < n_fts0) continue;
    Feature* ft = *it;
    const int i = index.at(ft->point);
    out[i].matched = 1; out[i].tried = 1; out[i].order = order++;
    out[i].search_level = ft->level;
    out[i].px[0] = ft->px[0]; out[i].px[1] = ft->px[1];
    out[i].align_ok = 1;
  }
  for (int i = 0; i < M; ++i)
    if (pts[i] && (pts[i]->n_failed_reproj_ > 0 || pts[i]->n_succeeded_reproj_ > 0)) out[i].tried = 1;
  // the features the walk created point at Points that die with this call: take them out of the frame again
  while (frame->fts_.size() > n_fts0) { delete frame->fts_.back(); frame->fts_.pop_back(); }
}

// ---- a13b: the seed stage of Reprojector::reprojectMap through the reference's own member functions -------------------------------------------
// Seeds as orc_seed_obs (Seed{ftr = the feature in keyframe kf_handles[ref_frame], mu, sigma2}). Per seed, in list order, Reprojector::reprojectorSeed
// (:531-552) — where reprojectMap's loop over depth_filter_->seeds_ (:312-317, host side: it selects the seeds) calls it; then lines :319-327
// restated verbatim around the reference's reprojectorSeeds (:431-503: per-cell stable sort by sigma2, first seed findMatchSeed accepts, the
// TYPE_TEMPORARY point and the feature it creates). n_matches_in = n_matches_ when the stage starts.
void ref_reproject_seeds(void* cur_handle, int n_kf, void* const* kf_handles, int S, const orc_seed_obs* seeds, const orc_reproj_grid* grid,
                         const int32_t* cell_order, int n_matches_in, orc_reproj_result* out, orc_reproj_summary* summary) {
  FrameHandle* cur = (FrameHandle*)cur_handle;
  FramePtr frame = cur->frame;
  std::memset(summary, 0, sizeof *summary);
  for (int i = 0; i < S; ++i) { std::memset(&out[i], 0, sizeof out[i]); out[i].order = -1; out[i].cell = -1; }
  Config::maxFts() = (size_t)grid->max_fts;
  Map map;
  Reprojector rep(frame->cam_, map);
  if (rep.grid_.cell_size != grid->cell_size || rep.grid_.grid_n_cols != grid->n_cols || rep.grid_.grid_n_rows != grid->n_rows) { summary->n_matches = -1; return; }
  rep.resetGrid();
  rep.grid_.cell_order.assign(cell_order, cell_order + rep.grid_.cells.size());
  rep.matcher_.options_.align_max_iter = grid->align_max_iter;
  rep.n_matches_ = (size_t)n_matches_in;
  std::vector<std::unique_ptr<Feature>> feats(S);
  std::list<Seed> list;  // DepthFilter::seeds_
  std::map<Feature*, int> index;
  for (int i = 0; i < S; ++i) {
    const orc_seed_obs& s = seeds[i];
    if (s.ref_frame < 0 || s.ref_frame >= n_kf) continue;
    Frame* kf = ((FrameHandle*)kf_handles[s.ref_frame])->frame.get();
    feats[i].reset(new Feature(kf, Vector2d(s.px[0], s.px[1]), Vector3d(s.f[0], s.f[1], s.f[2]), s.level));
    feats[i]->type = (Feature::FeatureType)s.ftr_type;
    feats[i]->grad = Vector2d(s.grad[0], s.grad[1]);
    list.emplace_back(feats[i].get(), 1.0f, 1.0f, 1.0f);
    Seed& seed = list.back();
    seed.mu = s.mu; seed.sigma2 = s.sigma2;
    index[feats[i].get()] = i;
    auto it = std::prev(list.end());
    if (rep.reprojectorSeed(frame, seed, it)) {
      out[i].in_frame = 1;
      summary->n_in_frame++;
      // the pixel and cell reprojectorSeed computed: the candidate it just appended
      for (size_t k = 0; k < rep.grid_.seeds.size(); ++k)
        if (!rep.grid_.seeds[k]->empty() && &rep.grid_.seeds[k]->back().seed == &seed) {
          out[i].cell = (int32_t)k;
          out[i].px[0] = rep.grid_.seeds[k]->back().px[0]; out[i].px[1] = rep.grid_.seeds[k]->back().px[1];
        }
    }
  }
  const size_t n_fts0 = frame->fts_.size();
  // ---- src/reprojector.cpp:319-327 ----
  for (size_t i = 0; i < rep.grid_.seeds.size(); ++i) {
    if (rep.reprojectorSeeds(*rep.grid_.seeds.at(rep.grid_.cell_order[i]), frame)) ++rep.n_matches_;
    if (rep.n_matches_ >= (size_t)Config::maxFts()) break;
  }
  summary->n_matches = (int32_t)rep.n_matches_;
  int order = 0;
  size_t k = 0;
  for (auto it = frame->fts_.begin(); it != frame->fts_.end(); ++it, ++k) {
    if (k < n_fts0) continue;
    Feature* ft = *it;
    const int i = index.at(ft->point->hostFeature_);
    out[i].matched = 1; out[i].tried = 1; out[i].order = order++;
    out[i].search_level = ft->level;
    out[i].px[0] = ft->px[0]; out[i].px[1] = ft->px[1];
    out[i].align_ok = 1;
  }
  while (frame->fts_.size() > n_fts0) { delete frame->fts_.back(); frame->fts_.pop_back(); }  // their TYPE_TEMPORARY points stay with `map` and die with it
}

// ---- N3: DepthFilter::observeDepthRow (src/depth_filter.cpp:580-675) — the reference's own function, one seed per call so that the outcome code of
// doLineStereo can be read off the RunningStats counters. Seeds as orc_seed_obs (Seed{ftr = the feature in keyframe kf_handles[ref_frame], mu, sigma2}),
// results as orc_seed_result. The filter object is constructed once (its constructor starts the IndexThreadReduce workers, which stay idle); the
// mapping thread is never started.
void ref_depth_observe(void* cur_handle, int n_kf, void* const* kf_handles, double px_error_angle, int S, const orc_seed_obs* seeds, orc_seed_result* out) {
  FrameHandle* cur = (FrameHandle*)cur_handle;
  static DepthFilter* df = nullptr;
  if (!df) df = new DepthFilter(nullptr, DepthFilter::callback_t());
  df->px_error_angle_ = px_error_angle;
  df->active_frame_ = cur->frame;
  df->seeds_updating_halt_ = false;
  for (int i = 0; i < S; ++i) {
    const orc_seed_obs& s = seeds[i];
    orc_seed_result& o = out[i];
    std::memset(&o, 0, sizeof o);
    if (s.ref_frame < 0 || s.ref_frame >= n_kf) continue;
    Frame* kf = ((FrameHandle*)kf_handles[s.ref_frame])->frame.get();
    Feature obs(kf, Vector2d(s.px[0], s.px[1]), Vector3d(s.f[0], s.f[1], s.f[2]), s.level);
    obs.type = (Feature::FeatureType)s.ftr_type;
    obs.grad = Vector2d(s.grad[0], s.grad[1]);
    df->seeds_.clear();
    df->seeds_.emplace_back(&obs, 1.0f, 1.0f, 1.0f);   // Seed(ftr, depth_mean, depth_min, converge_threshold): only ftr, mu, sigma2, b are read below
    Seed& seed = df->seeds_.back();
    seed.mu = s.mu; seed.sigma2 = s.sigma2; seed.is_update = false;
    const float b0 = seed.b;
    RunningStats st;
    df->observeDepthRow(0, 1, &st);
    o.is_update = seed.is_update ? 1 : 0;
    o.is_valid = seed.isValid ? 1 : 0;
    o.res = st.n_updates ? 1 : st.n_fail_lsd ? -1 : st.n_fail_triangulation ? -2 : st.n_fail_alignment ? -3 : st.n_fail_score ? -4 : (seed.b != b0 ? -5 : 0);
    o.epl_start[0] = seed.eplStart[0]; o.epl_start[1] = seed.eplStart[1]; o.epl_end[0] = seed.eplEnd[0]; o.epl_end[1] = seed.eplEnd[1];
    o.mu = seed.mu; o.sigma2 = seed.sigma2;
    if (st.n_updates) {
      o.z = (double)seed.vec_distance.back();  // 1 / mu after the update (float); the triangulated z itself is a local of observeDepthRow
      o.px_cur[0] = seed.last_matched_px[0]; o.px_cur[1] = seed.last_matched_px[1];
      o.search_level = seed.last_matched_level;
    }
    df->seeds_.clear();
  }
  df->active_frame_.reset();
}

// N2: hso::shiTomasiScore (src/vikit/vision.cpp:111-151) of the reference for a list of pixels
void ref_shi_tomasi(const uint8_t* img, int w, int h, int n, const int32_t* xy, float* out) {
  const cv::Mat m = mat_from_u8(img, w, h);
  for (int i = 0; i < n; ++i) out[i] = hso::shiTomasiScore(m, xy[2 * i], xy[2 * i + 1]);
}

int ref_check_ncc(const float* p1, const float* p2, float thresh) {
  Matcher m;
  return m.checkNCC(const_cast<float*>(p1), const_cast<float*>(p2), thresh) ? 1 : 0;
}

// ---- a16: pose_optimizer::optimizeLevenbergMarquardt3rd (src/pose_optimizer.cpp:399-771) on the flattened inputs of orc_pose_optimize -------
// frame_handle: any reference Frame of the camera (its image is irrelevant to the optimiser); host_handles: K frames whose poses are set from
// T_host_w. Features without a point are appended so that frame->fts_.size() == n_fts_total (:696).
void ref_pose_optimize(void* frame_handle, int K, void* const* host_handles, double reproj_thresh, int n_iter, int n_fts_total, int F, const double* f,
                       const double* p_host, const int32_t* host_idx, const double* T_host_w, const double* grad, const int8_t* level,
                       const int8_t* ftype, const int8_t* ptype, const double T_f_w_in[12], uint8_t* outlier_out, orc_pose_result* out) {
  FrameHandle* fh = (FrameHandle*)frame_handle;
  clear_features(fh);
  Frame* fr = fh->frame.get();
  fr->T_f_w_ = se3_from_rt(T_f_w_in);
  fr->Cov_.setZero();
  fr->m_error_in_px = 1.f;
  for (int k = 0; k < K; ++k) ((FrameHandle*)host_handles[k])->frame->T_f_w_ = se3_from_rt(T_host_w + 12 * k);
  std::vector<std::unique_ptr<Feature>> host_fts;
  std::vector<Feature*> fts;
  for (int i = 0; i < F; ++i) {
    Frame* hf = ((FrameHandle*)host_handles[host_idx[i]])->frame.get();
    const Vector3d ph(p_host[3 * i], p_host[3 * i + 1], p_host[3 * i + 2]);
    // pHost = hostFeature_->f * (1.0 / idist_) (:431): f = p_host, idist = 1 reproduces p_host exactly
    host_fts.emplace_back(new Feature(hf, Vector2d(0, 0), ph, 0));
    Point* pt = new Point(Vector3d(0, 0, 0), host_fts.back().get());
    pt->hostFeature_ = host_fts.back().get();
    pt->idist_ = 1.0;
    pt->type_ = (Point::PointType)ptype[i];
    fh->points.push_back(pt);
    Feature* ft = new Feature(fr, pt, Vector2d(0, 0), Vector3d(f[3 * i], f[3 * i + 1], f[3 * i + 2]), level[i]);
    ft->type = (Feature::FeatureType)ftype[i];
    ft->grad = Vector2d(grad[2 * i], grad[2 * i + 1]);
    fr->fts_.push_back(ft);
    fts.push_back(ft);
  }
  for (int i = F; i < n_fts_total; ++i) fr->fts_.push_back(new Feature(fr, Vector2d(0, 0), Vector3d(0, 0, 1), 0));
  double scale = 0, e_init = 0, e_final = 0;
  size_t num_obs = (size_t)F;  // the caller passes the number of observations in (frame_handler_mono.cpp:239-243)
  pose_optimizer::optimizeLevenbergMarquardt3rd(reproj_thresh, (size_t)n_iter, false, fh->frame, scale, e_init, e_final, num_obs);
  std::memset(out, 0, sizeof *out);
  se3_to_rt(fr->T_f_w_, out->T_f_w);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) out->cov[6 * i + j] = fr->Cov_(i, j);
  out->estimated_scale = scale; out->error_init = e_init; out->error_final = e_final;
  out->num_obs = (uint64_t)num_obs;
  out->error_in_px = fr->m_error_in_px;
  out->n_trials_total = -1;  // not observable from outside
  for (int i = 0; i < F; ++i) outlier_out[i] = fts[i]->point == NULL ? 1 : 0;
  clear_features(fh);
}

}  // extern "C"
