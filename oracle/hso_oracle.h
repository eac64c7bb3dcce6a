/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * C interface of the CPU restatement of HSO's per-frame tracking hot path. Every function cites the
 * reference file:line it follows (paths relative to /root/reference). Nothing under hso_b200/ may call
 * this library; it exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * arm can check and time the reference algorithm.
 *
 * Pose layout everywhere: 12 doubles, row-major 3x4 [R | t].
 */
#ifndef HSO_ORACLE_H
#define HSO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Camera models of src/camera.cpp. model: 0 = PinholeCamera (radtan when |d0|>1e-7, camera.cpp:30-38,99-125),
 * 1 = FOVCamera (atan/omega in d[0]; undistort!=0 => plain pinhole, camera.cpp:199-221),
 * 2 = EquidistantCamera (always undistorted up-front => plain pinhole, camera.cpp:307-315). */
typedef struct {
  int model, width, height, undistort;
  double fx, fy, cx, cy;
  double d[5];
} orc_cam;

/* ---- a12: Sophus SE3 ---- */
void orc_se3_exp(const double tangent[6], double rt_out[12]);
void orc_se3_log(const double rt[12], double tangent_out[6]);
void orc_se3_mul(const double a[12], const double b[12], double out[12]);
void orc_se3_inverse(const double a[12], double out[12]);
/* quaternion-level API used by the Sophus known-answer tests: q = (w,x,y,z), t */
void orc_se3_from_qt(const double q[4], const double t[3], double rt_out[12]);

/* ---- LDLT (Eigen::LDLT restatement) ---- */
void orc_ldlt_solve7(const double* A, const double* b, double* x);
void orc_ldlt_solve6(const double* A, const double* b, double* x);

/* ---- a11: world2cam ---- */
void orc_world2cam(const orc_cam* cam, const double xyz[3], double px_out[2]);

/* ---- a1: pyramid (src/frame.cpp:296-314, src/vikit/vision.cpp:19-44,70-108) ----
 * mode: -1 = what the reference does on x86 (SSE2 rounding iff in.cols%16==0), 0 = scalar truncating, 1 = SSE2 rounding */
void orc_half_sample(const uint8_t* in, int w, int h, uint8_t* out, int mode);
/* Builds levels 1..n_levels-1 into `out` (concatenated, tightly packed); level sizes written to lw/lh[0..n_levels).
 * Returns 0 on the halfSample path, 1 if the cv::resize(INTER_LINEAR) path was taken (W or H not a multiple of 16). */
int orc_create_pyramid(const uint8_t* img, int W, int H, int n_levels, uint8_t* out, int* lw, int* lh);
/* cv::resize(src, dst, Size(dw,dh), 0, 0, INTER_LINEAR) for CV_8UC1 — OpenCV's published fixed-point algorithm
 * (modules/imgproc/src/resize.cpp, INTER_RESIZE_COEF_BITS=11); pinned against cv2 4.13 golden vectors. */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);

/* ---- a2: Sobel 5x5 CV_16S BORDER_REPLICATE + interior statistics (src/frame.cpp:205-246) ---- */
void orc_sobel5(const uint8_t* img, int w, int h, int16_t* gx, int16_t* gy);
void orc_frame_stats(const uint8_t* img, const int16_t* gx, const int16_t* gy, int w, int h, float* integral, float* grad_mean);

/* ---- a4: makeDepthRef (src/CoarseTracker.cpp:210-240) ----
 * has_point[i]==0 => -1 ; T_ref_host[i] = ref.T_f_w * host.T_f_w^-1 is formed here from the two world poses. */
void orc_accumulator7(int n, const float* J /*7n*/, const float* w /*n*/, float* H49);
void orc_make_depth_ref(const double T_ref_w[12], int F, const uint8_t* has_point, const double* f_host /*3F*/,
                        const double* idist /*F*/, const double* T_host_w /*12F*/, double* dist_out /*F*/);

/* ---- a3,a5-a10: CoarseTracker ---- */
typedef struct {
  int level, iter;              /* iter = -1: evaluation at level entry (computeResiduals+computeGS before the loop) */
  double T_eval[12];            /* pose the residuals were evaluated at (trial pose for iter>=0) */
  float a_eval;                 /* exposure ratio evaluated */
  float lambda;                 /* damping used to solve for this trial (iter>=0) */
  double H[49], b[7];           /* system solved for this trial (iter>=0) / system built at entry (iter=-1) */
  double step[7];               /* (extrapolated, NaN-guarded) step ; zeros for iter=-1 */
  double energy;                /* E/total_terms of this evaluation */
  int total_terms, saturated_terms;
  int accepted;                 /* iter>=0: energy_new < energy_old */
  float huber, outlier;         /* thresholds of the level */
} orc_trace;

typedef struct {
  int inverse_comp, max_level, min_level, n_iter;
} orc_track_params;

/* Full CoarseTracker::run (src/CoarseTracker.cpp:51-208) on flattened inputs.
 * levels: ref_levels[l]/cur_levels[l] point to tightly packed u8 images of lw[l] x lh[l].
 * px: level-0 pixel (2F doubles), f: unit bearing (3F), dist: from makeDepthRef (<0: feature skipped).
 * T_cur_ref_io / a_io: initial guess in, result out. trace (optional) receives up to trace_cap entries.
 * Returns the reference's return value size_t(float(total_terms)/PATCH_AREA) (CoarseTracker.cpp:207). */
uint64_t orc_coarse_track(const orc_cam* cam, const orc_track_params* prm, int n_levels,
                          const uint8_t* const* ref_levels, const uint8_t* const* cur_levels, const int* lw, const int* lh,
                          int F, const double* px, const double* f, const double* dist,
                          double T_cur_ref_io[12], float* a_io,
                          orc_trace* trace, int trace_cap, int* trace_len, int* n_evals_out);

/* One residual evaluation + normal equations at a given state (computeResiduals + computeGS,
 * src/CoarseTracker.cpp:242-414,499-525) with thresholds supplied — used for per-iteration parity replay. */
void orc_track_eval(const orc_cam* cam, int inverse_comp, int level, int max_level,
                    const uint8_t* ref_img, const uint8_t* cur_img, int w, int h,
                    int F, const double* px, const double* f, const double* dist,
                    const double T[12], float a, float huber, float outlier,
                    double H_out[49], double b_out[7], double* energy_out, int* total_terms, int* saturated_terms);
/* selectRobustFunctionLevel (src/CoarseTracker.cpp:530-644). */
void orc_track_select_robust(const orc_cam* cam, int level, int max_level, const uint8_t* ref_img, const uint8_t* cur_img,
                             int w, int h, int F, const double* px, const double* f, const double* dist,
                             const double T[12], float a, float* huber_out, float* outlier_out, int* n_errors);
/* Damped solve + extrapolation + NaN guard (src/CoarseTracker.cpp:112-124). */
void orc_track_solve(const double H[49], const double b[7], float lambda, double step_out[7]);

/* ---- a13-a15: matcher / feature_alignment ---- */
int orc_align2d(const uint8_t* cur_img, int cols, int rows, int stride, const float* ref_patch_with_border /*100*/,
                const float* ref_patch /*64*/, int n_iter, double px_io[2], float* cur_patch_out /*64 or NULL*/);
int orc_align1d(const uint8_t* cur_img, int cols, int rows, int stride, const float dir[2],
                const float* ref_patch_with_border, const float* ref_patch, int n_iter, double px_io[2],
                double* h_inv_out, float* cur_patch_out);
/* warp::warpAffine float overload (src/matcher.cpp:120-155), halfpatch = 5 => 10x10 */
void orc_warp_affine(const double A_cur_ref[4], const uint8_t* img_ref, int cols, int rows, const double px_ref[2],
                     int level_ref, int search_level, int halfpatch_size, float* patch_out);
int orc_get_best_search_level(const double A_cur_ref[4], int max_level);
int orc_check_ncc(const float* p1, const float* p2, float thresh);
int orc_check_normal(const int16_t* sobx, const int16_t* soby, int cols, const double px_level[2], const double normal[2], float thresh);

/* Job/result records shared in layout with include/hso_b200.h's hso_align_job / hso_align_result. */
typedef struct {
  int32_t ref_level;        /* pyramid level of the reference observation (Feature::level) */
  int32_t search_level;     /* from getBestSearchLevel */
  int32_t type;             /* 0 corner, 1 edgelet, 2 gradient (feature.h:36) */
  int32_t scale_patch;      /* !=0: multiply the warped patch by exposure_rat (matcher.cpp:317-330) */
  double px_ref[2];         /* ref feature pixel, level 0 */
  double A_cur_ref[4];      /* row-major 2x2 */
  double grad[2];           /* ref feature gradient direction (edgelets) */
  double px_cur[2];         /* initial estimate, level 0 pixels */
  float exposure_rat;
  float ncc_thresh;         /* checkNCC threshold; 0 = findMatchDirect's 0.7 (matcher.cpp:366); findMatchSeed: 0.8 (:510) */
} orc_align_job;
typedef struct {
  int32_t ok;               /* findMatchDirect's return value */
  int32_t align_converged;  /* result of align1D/align2D alone */
  double px_cur[2];         /* written even on failure (matcher.cpp:372) */
  double h_inv;             /* align1D only */
} orc_align_result;
/* The post-getWarpMatrixAffine part of Matcher::findMatchDirect (src/matcher.cpp:310-375). sobel may be NULL when
 * no edgelet jobs are present; arrays are per pyramid level 0..2. */
void orc_match_direct_batch(int M, const orc_align_job* jobs, const uint8_t* const* ref_levels, const uint8_t* const* cur_levels,
                            const int* lw, const int* lh, const int16_t* const* cur_sobx, const int16_t* const* cur_soby,
                            int align_max_iter, orc_align_result* out);

/* ---- N1: Reprojector::reprojectMap data path (src/reprojector.cpp:88-331,504-615) + whole Matcher::findMatchDirect ----
 * Record layouts shared with include/hso_b200.h (hso_reproj_*); see there for the field meaning. */
typedef struct {
  double p_host[3], px_ref[2], f_ref[3], grad[2], depth_ref;
  int32_t host_pose, ref_pose, ref_frame, ref_level, ftr_type, pt_type, pt_ftr_type, scale_patch;
  float exposure_rat, pad_;
} orc_reproj_cand;
typedef struct { int32_t cell_size, n_cols, n_rows, max_fts, align_max_iter, pad_; } orc_reproj_grid;
typedef struct {
  int32_t in_frame, cell, tried, matched, search_level, order, align_ok, pad_;
  double px[2], A_cur_ref[4];
} orc_reproj_result;
typedef struct { int32_t n_in_frame, n_matches, n_trials, used_cell_all; } orc_reproj_summary;
/* cam2world of the three camera models (src/camera.cpp:66-87,169-190,297-300); unit-norm bearing. */
void orc_cam2world(const orc_cam* cam, double u, double v, double xyz_out[3]);
void orc_cv_undistort_point(const float K[4] /*fx fy cx cy*/, const float D[5], float u, float v, float out[2]);
/* warp::getWarpMatrixAffine (src/matcher.cpp:46-72); A row-major. */
void orc_get_warp_matrix_affine(const orc_cam* cam, const double px_ref[2], const double f_ref[3], double depth_ref, const double T_cur_ref[12],
                                int level_ref, double A_cur_ref[4]);
/* ref_levels[frame][level]: pyramids of the frames hso_reproj_cand::ref_frame indexes. max_search_level = Config::nPyrLevels()-1. */
void orc_reproject_match(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int M, const orc_reproj_cand* cands,
                         const orc_reproj_grid* grid, const int32_t* cell_order, int max_search_level,
                         const uint8_t* const* const* ref_levels, const uint8_t* const* cur_levels, const int* lw, const int* lh,
                         const int16_t* const* cur_sobx, const int16_t* const* cur_soby, orc_reproj_result* out, orc_reproj_summary* summary);
void orc_reproject_speculative(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int M, const orc_reproj_cand* cands,
                               const orc_reproj_grid* grid, int max_search_level, const uint8_t* const* const* ref_levels,
                               const uint8_t* const* cur_levels, const int* lw, const int* lh, const int16_t* const* cur_sobx,
                               const int16_t* const* cur_soby, orc_reproj_result* out, double* px_after);
/* The selection walk alone (cells, ordering, three passes / reprojectCellAll) with findMatchDirect's outcome supplied per candidate:
 * io[i].in_frame / cell are inputs, tried / matched / order are written. */
void orc_reproject_select(int M, const orc_reproj_cand* cands, const uint8_t* match_ok, const orc_reproj_grid* grid, const int32_t* cell_order,
                          orc_reproj_result* io, orc_reproj_summary* summary);

/* ---- a16,a17: pose_optimizer::optimizeLevenbergMarquardt3rd (src/pose_optimizer.cpp:399-771) ---- */
typedef struct {
  double T_f_w[12];       /* optimised pose */
  double cov[36];         /* Frame::Cov_ */
  double estimated_scale, error_init, error_final;
  uint64_t num_obs;
  float error_in_px;      /* Frame::m_error_in_px */
  int n_trials_total;     /* number of (build, solve, re-evaluate) trials executed */
  int early_return;       /* 1: no observations (pose_optimizer.cpp:456) */
} orc_pose_result;
/* Flattened inputs: per feature with a point — f (3), pHost = f_host/idist (3), host index into T_host_w (K x 12),
 * grad (2), level, ftype (0/1/2), ptype (Point::PointType, 1 = TEMPORARY). n_fts_total = frame->fts_.size()
 * including features without point (decides the <80 threshold, pose_optimizer.cpp:696). */
void orc_pose_optimize(double reproj_thresh, int n_iter, double err_mult2 /* cam errorMultiplier2 */, int n_fts_total,
                       int F, const double* f, const double* p_host, const int32_t* host_idx, const double* T_host_w,
                       const double* grad, const int8_t* level, const int8_t* ftype, const int8_t* ptype,
                       const double T_f_w_in[12], uint8_t* outlier_out /*F*/, orc_pose_result* out);

/* ---- N3: DepthFilter::observeDepthRow (src/depth_filter.cpp:580-675) + Matcher::doLineStereo (src/matcher.cpp:802-1049) ----
 * Record layouts shared with include/hso_b200.h (hso_seed_obs / hso_seed_result). */
typedef struct {
  double px[2], f[3], grad[2];
  int32_t ref_frame, ref_pose, level, ftr_type;
  float mu, sigma2, exposure_rat, pad_;
} orc_seed_obs;
typedef struct {
  int32_t is_update, is_valid, res, search_level;
  int32_t epl_start[2], epl_end[2];
  float mu, sigma2;
  double z;
  double px_cur[2];
} orc_seed_result;
void orc_depth_observe(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, double px_error_angle, int S,
                       const orc_seed_obs* seeds, int max_search_level, int align_max_iter, const uint8_t* const* const* ref_levels,
                       const uint8_t* const* cur_levels, const int* lw, const int* lh, const int16_t* const* cur_sobx, const int16_t* const* cur_soby,
                       orc_seed_result* out);

/* ---- a13b: the seed stage of Reprojector::reprojectMap (src/reprojector.cpp:309-328,431-503,531-552) + Matcher::findMatchSeed
 * (src/matcher.cpp:442-518). Seeds as orc_seed_obs, results as orc_reproj_result (see include/hso_b200.h, hso_reproject_seeds). */
void orc_reproject_seeds(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int S, const orc_seed_obs* seeds,
                         const orc_reproj_grid* grid, const int32_t* cell_order, int n_matches_in, int max_search_level,
                         const uint8_t* const* const* ref_levels, const uint8_t* const* cur_levels, const int* lw, const int* lh,
                         const int16_t* const* cur_sobx, const int16_t* const* cur_soby, orc_reproj_result* out, orc_reproj_summary* summary);
/* findMatchSeed for EVERY seed that entered a cell (what the CUDA path computes speculatively); px_after: the pixel findMatchSeed leaves. */
void orc_reproject_seeds_speculative(const orc_cam* cam, const double T_cur_w[12], int n_poses, const double* T_f_w, int S, const orc_seed_obs* seeds,
                                     const orc_reproj_grid* grid, int max_search_level, const uint8_t* const* const* ref_levels,
                                     const uint8_t* const* cur_levels, const int* lw, const int* lh, const int16_t* const* cur_sobx,
                                     const int16_t* const* cur_soby, orc_reproj_result* out, double* px_after);
/* The walk alone (per-cell stable sort by sigma2, first accepted seed per cell, break at maxFts) with findMatchSeed's outcome supplied. */
void orc_seed_select(int S, const orc_seed_obs* seeds, const uint8_t* match_ok, const orc_reproj_grid* grid, const int32_t* cell_order, int n_matches_in,
                     orc_reproj_result* io, orc_reproj_summary* summary);

/* ---- N4: undistortion maps (src/camera.cpp:47-54,223-265,317-363) and cv::remap INTER_LINEAR (:127-131,267-271,365-369) ---- */
void orc_convert_maps(const float* mapx, const float* mapy, int n, int16_t* map1, uint16_t* map2);
int orc_init_undistort_maps(const orc_cam* cam, int16_t* map1 /*[h][w][2]*/, uint16_t* map2 /*[h][w]*/);
void orc_remap_linear_u8(const uint8_t* src, int sw, int sh, int sstride, const int16_t* map1, const uint16_t* map2, int dw, int dh, uint8_t* dst);

/* ---- N2: FeatureExtractor::fastDetectST (src/feature_detection.cpp:498-523; thirdparty/fast) ---- */
typedef struct {
  int16_t x, y;       /* level pixel */
  int32_t score;      /* fast_corner_score_9 */
  float shi_tomasi;   /* hso::shiTomasiScore (src/vikit/vision.cpp:111-151) */
} orc_corner;
int orc_fast9_corners(const uint8_t* img, int w, int h, int stride, int threshold, int16_t* xy, int32_t* scores, int cap);
int orc_fast_detect(const uint8_t* img, int w, int h, int stride, int threshold, int border, orc_corner* out, int cap);

#ifdef __cplusplus
}
#endif
#endif
