/* hso_b200 — C-ABI of the B200-native (sm_100a CUDA) implementation of HSO's per-frame tracking hot path.
 *
 * The reference (luodongting/HSO) has no plugin/FFI surface: its hot path is reached through concrete C++
 * classes with Eigen/Sophus/cv::Mat types. This header is the boundary a replacement sits
 * under. Each entry point names the reference interface it replaces (paths relative to the reference root).
 * The C++ host shim in hso_b200/host/ keeps the reference's class/function names on top of these calls;
 * INTEGRATION.md shows the glue a maintainer adds inside the reference tree.
 *
 * Conventions: POD only, caller-owned host buffers, row-major; poses are 12 doubles = 3x4 [R | t];
 * every call returns HSO_OK (0) or a negative hso_status; no exceptions/aborts cross the boundary;
 * one hso_ctx per host thread (a ctx owns one CUDA stream; calls on one ctx are not thread-safe,
 * distinct contexts are independent). There is NO CPU fallback: without a CUDA device hso_create fails.
 */
#ifndef HSO_B200_H
#define HSO_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HSO_OK = 0,
  HSO_ERR_INVALID = -1,     /* bad argument (mirrors the reference's size/type check, src/frame.cpp:85-86) */
  HSO_ERR_CUDA = -2,        /* CUDA runtime error; see hso_last_error */
  HSO_ERR_NO_DEVICE = -3,   /* no usable sm_100 device */
  HSO_ERR_CAPACITY = -4,    /* frame table / feature capacity exceeded */
  HSO_ERR_BAD_FRAME = -5    /* unknown or released frame id */
} hso_status;

typedef struct hso_ctx hso_ctx;
typedef int32_t hso_frame_id;

/* Camera models of src/camera.cpp. model 0 = PinholeCamera (radial-tangential distortion applied inside world2cam
 * when |d[0]| > 1e-7, camera.cpp:36,99-125), 1 = FOVCamera (omega in d[0]; undistort != 0 => plain pinhole,
 * camera.cpp:199-221), 2 = EquidistantCamera (image undistorted up-front => plain pinhole, camera.cpp:307-315). */
typedef struct {
  int32_t model, width, height, undistort;
  double fx, fy, cx, cy;
  double d[5];
} hso_cam;

/* Subset of hso::Config that the path reads (src/config.cpp:28-64). Zero-initialise then call hso_cfg_default. */
typedef struct {
  int32_t n_pyr_levels;      /* Config::nPyrLevels = 3 : levels that get Sobel images / are searched by the matcher */
  int32_t klt_max_level;     /* Config::kltMaxLevel = 4 : pyramid has max(n_pyr_levels, klt_max_level+1) levels */
  int32_t max_frames;        /* frames resident on the device per context (ring; default 16) */
  int32_t max_features;      /* per-job feature capacity (default 8192) */
  int32_t materialize_sobel; /* 1: keep the int16 Sobel L0..L2 images on device (src/frame.cpp:216-220) */
  int32_t reserved[3];
} hso_cfg;
void hso_cfg_default(hso_cfg* cfg);

/* ---- lifetime ------------------------------------------------------------------------------------------ */
int hso_create(int device, const hso_cam* cam, const hso_cfg* cfg, hso_ctx** out);
void hso_destroy(hso_ctx* ctx);
const char* hso_last_error(const hso_ctx* ctx);
/* Use an existing CUDA stream (cudaStream_t as void*) for all work of this context; NULL restores the ctx's own. */
int hso_set_stream(hso_ctx* ctx, void* cuda_stream);
void* hso_get_stream(hso_ctx* ctx);
int hso_synchronize(hso_ctx* ctx);
/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t hso_kernel_launches(const hso_ctx* ctx);

/* ---- F1: Frame construction — replaces hso::Frame::Frame / initFrame (include/hso/frame.h:131, src/frame.cpp:82-96):
 * frame_utils::createImgPyramid (src/frame.cpp:296-314, halfSample src/vikit/vision.cpp:19-108) and
 * Frame::prepareForFeatureDetect (src/frame.cpp:205-246). The pyramid stays on the device. -------------------------- */
/* img: CV_8UC1 W x H with `stride` bytes per row. W,H must equal the camera's (else HSO_ERR_INVALID, as frame.cpp:85).
 * integral/grad_mean receive Frame::integralImage_ / Frame::gradMean_ (may be NULL to skip the read-back). */
int hso_frame_upload(hso_ctx* ctx, const uint8_t* img, int W, int H, int stride, hso_frame_id* out, float* integral, float* grad_mean);
/* Batched form: B images -> B frames with one synchronisation. imgs[i] are host pointers (pinned or pageable). */
int hso_frame_upload_batch(hso_ctx* ctx, int B, const uint8_t* const* imgs, int W, int H, int stride, hso_frame_id* out,
                           float* integral, float* grad_mean);
/* Device-resident variant for benchmarking: raw level-0 images already live in device memory (dev_imgs = B device ptrs). */
int hso_frame_build_batch_device(hso_ctx* ctx, int B, const void* const* dev_imgs, int W, int H, int stride, hso_frame_id* out);
/* Rebuild existing frames `ids` from device-resident images (asynchronous on the ctx stream, no allocation). */
int hso_frame_rebuild_batch_device(hso_ctx* ctx, int B, const void* const* dev_imgs, int W, int H, int stride, const hso_frame_id* ids);
int hso_frame_stats(hso_ctx* ctx, hso_frame_id id, float* integral, float* grad_mean);
int hso_frame_level_size(hso_ctx* ctx, hso_frame_id id, int level, int* w, int* h);
/* Host mirrors for consumers that stay on the CPU (mapping thread, feature detection). */
int hso_frame_download_level(hso_ctx* ctx, hso_frame_id id, int level, uint8_t* dst);
int hso_frame_download_sobel(hso_ctx* ctx, hso_frame_id id, int level, int16_t* gx, int16_t* gy);
int hso_frame_release(hso_ctx* ctx, hso_frame_id id);
int hso_frame_release_batch(hso_ctx* ctx, int n, const hso_frame_id* ids);  /* ~Frame for n frames; returns the last error, releases the rest */
/* ---- N4 (next row): input side — what test/test_dataset.cpp:262-283 does to a raw image before addImage: ImageReader::readImage's
 * cv::resize(image, image, m_img_new_size) (src/ImageReader.cpp:80) when (raw_w, raw_h) differs from the camera size, then
 * cam->undistortImage(image, image) = cv::remap(INTER_LINEAR) through the CV_16SC2 maps the camera constructor builds
 * (src/camera.cpp:47-54,127-131 pinhole; :223-271 FOV; :317-369 equidistant) when undistort != 0, then new Frame (pyramid + statistics).
 * Bit-exact against OpenCV (cv2 4.13 golden vectors). hso_undistort_maps returns the maps ([H][W][2] int16, [H][W] uint16) for host consumers. */
int hso_frame_upload_raw_batch(hso_ctx* ctx, int B, const uint8_t* const* imgs, int raw_w, int raw_h, int stride, int undistort, hso_frame_id* out,
                               float* integral, float* grad_mean);
int hso_undistort_maps(hso_ctx* ctx, int16_t* map1, uint16_t* map2);


/* ---- F2: sparse direct image alignment — replaces size_t hso::CoarseTracker::run(FramePtr ref, FramePtr cur)
 * (include/hso/CoarseTracker.h:134,141 ; src/CoarseTracker.cpp:51-208). ------------------------------------------------ */
typedef struct {
  int32_t inverse_comp;  /* ctor arg inverse_composition (frame_handler_mono.cpp:184-203) */
  int32_t max_level;     /* Config::kltMaxLevel() = 4 */
  int32_t min_level;     /* Config::kltMinLevel()+1 = 1 ; 0 when relocalising */
  int32_t n_iter;        /* 50 ; 15 when relocalising */
} hso_track_params;

typedef struct {
  hso_frame_id ref, cur;
  int32_t n_features;    /* ref_frame->fts_.size() */
  int32_t reserved;
  const double* px;      /* 2F: Feature::px (level-0 pixels) */
  const double* f;       /* 3F: Feature::f (unit bearing) */
  const double* dist;    /* F : makeDepthRef output, <0 = no point / behind camera (CoarseTracker.cpp:210-240) */
  double T_cur_ref[12];  /* in: cur.T_f_w * ref.T_f_w^-1 (CoarseTracker.cpp:63) */
  float exposure_rat;    /* in: cur.integralImage_/ref.integralImage_ (CoarseTracker.cpp:60); < 0: formed on the device from the two frames' statistics */
  float reserved2;
  /* Optional compact layout, 32 B per feature instead of 48 (the end-to-end call is bound by the host->device copies): when both pointers are
   * non-null, px / f / dist are ignored. xyz = f * dist exactly as the caller's gather loop over ref_frame->fts_ forms it — the reference's own
   * expression, Vector3d xyz_ref((*it_ft)->f*dist), src/CoarseTracker.cpp:292 — with the features that have no point / a negative distance left out
   * (the reference skips them in every stage, :290,433,455,557); px as float32, which is exact: the tracker only ever uses
   * (float)(px * 2^-level) (:437-441), and a power-of-two factor commutes with the rounding to float. Results are bit-identical to the wide layout.
   * All jobs of a batch use the same layout. */
  const double* xyz;     /* [n_features][3] */
  const float* px32;     /* [n_features][2] */
} hso_track_job;

typedef struct {
  double T_cur_ref[12];        /* out: m_T_cur_ref */
  float exposure_rat;          /* out: m_exposure_rat */
  int32_t n_iters;             /* LM trials executed over all levels */
  int32_t n_evals;             /* residual evaluations = n_iters + number of levels */
  int32_t iters_per_level[8];
  uint64_t n_tracked;          /* return value of run(): size_t(float(total_terms)/PATCH_AREA) (CoarseTracker.cpp:207) */
  uint64_t visible_patch_evals[8]; /* per level: sum over evaluations of patches that produced terms (roofline accounting) */
  int32_t trace_len;           /* trace entries written for this problem */
  int32_t reserved;
  uint64_t cycles[8];          /* diagnostics, SM clock cycles summed over levels: [0] kernel, [1] setup (staging, reference patches,
                                  robust thresholds), [2] serial control (solve + SE3 update), [3] reference staging + reference patches,
                                  [4] current staging + threshold residuals, [5] median select, [6] MAD select */
} hso_track_result;

/* Per-evaluation trace for parity checking (one entry per computeResiduals call, CoarseTracker.cpp:102,141).
 * iter = -1: evaluation at level entry; iter >= 0: LM trial `iter`. */
typedef struct {
  int32_t level, iter;
  double T_eval[12];           /* pose the residuals were evaluated at */
  float a_eval, lambda;        /* exposure ratio evaluated; damping used to solve for this trial (0 for iter = -1) */
  double H[49], b[7];          /* normal equations built AT this evaluation (computeGS of this state, whether or not accepted) */
  double step[7];              /* extrapolated, NaN-guarded step that led to T_eval (zeros for iter = -1) */
  double energy;               /* E / total_terms of this evaluation */
  int32_t total_terms, saturated_terms, accepted;
  float huber, outlier;        /* thresholds of the level (selectRobustFunctionLevel) */
} hso_trace;

/* makeDepthRef helper on the host side of the boundary is hso::host::makeDepthRef (hso_b200/host); here dist is an input. */
int hso_coarse_track(hso_ctx* ctx, const hso_track_params* prm, const hso_track_job* job, hso_track_result* out,
                     hso_trace* trace, int trace_cap, int* trace_len);
/* Many independent (ref,cur) problems in one launch (the many-sequence mode). trace is [B][trace_cap] or NULL. */
int hso_coarse_track_batch(hso_ctx* ctx, const hso_track_params* prm, int B, const hso_track_job* jobs, hso_track_result* out,
                           hso_trace* trace, int trace_cap, int* trace_len);
/* The same call split in three so a caller can keep inputs resident and time the device part alone:
 * stage = H2D of the feature arrays, run = kernel launches only (asynchronous on the ctx stream), collect = sync + D2H. */
int hso_track_stage(hso_ctx* ctx, const hso_track_params* prm, int B, const hso_track_job* jobs, int trace_cap);
int hso_track_restage_frames(hso_ctx* ctx, int B, const hso_frame_id* ref, const hso_frame_id* cur);
int hso_track_run(hso_ctx* ctx);
int hso_track_collect(hso_ctx* ctx, hso_track_result* out, hso_trace* trace, int* trace_len);
/* Kernel timing for the roofline report: when on, CUDA events on the ctx stream bracket every per-level tracker launch;
 * hso_track_level_profile returns the accumulated device time and launch count of the level's kernel since profiling was enabled. */
int hso_track_set_profile(hso_ctx* ctx, int on);
int hso_track_level_profile(hso_ctx* ctx, int level, double* ms_total, uint64_t* launches);
/* Tuning knob: CTAs cooperating on one problem through a thread-block cluster (1,2,4,8; 0 = auto). */
int hso_track_set_cluster(hso_ctx* ctx, int ctas_per_problem, int threads_per_cta);
/* Same, for one pyramid level only (overrides hso_track_set_cluster for that level; 0,0 restores auto). */
int hso_track_set_level_shape(hso_ctx* ctx, int level, int ctas_per_problem, int threads_per_cta);
/* Launch shape the last run used for `level`: CTAs per problem, threads per CTA, mode (0 = images and caches in global memory, 1 = current level
 * + reference-patch cache in shared memory, 2 = both levels in shared memory, 3 = current level in shared memory + reference-patch cache streamed from L2 through a double-buffered
 * per-warp ring, 4 = as 3 with a single-buffered ring), and whether the |r| scratch of the threshold selection sat in
 * shared memory. The parity tests use it to prove that they exercise the shape the benchmark runs. */
int hso_track_get_level_shape(hso_ctx* ctx, int level, int* ctas, int* threads, int* mode, int* absres_smem);
/* Inverse-compositional mode: 1 (default) keeps both pyramid levels in shared memory and recomputes the reference samples per evaluation
 * whenever two copies of the level fit; 0 forces the cached-reference-patch path. Results are identical. */
int hso_track_set_ic_dual(hso_ctx* ctx, int enable);
/* Forward mode: where the reference-patch cache (precomputeReferencePatches, src/CoarseTracker.cpp:416-497) of a level lives. 0 (default):
 * in shared memory when image + cache fit one CTA, else in global memory streamed through a per-warp shared-memory ring (mode 3), else split
 * over a cluster; 1: mode 3 wherever it fits; -1: never mode 3. Results are identical for a given CTA count / thread count. */
int hso_track_set_stream_cache(hso_ctx* ctx, int mode);

/* F1 + F2 in one call for B independent streams — the front end of FrameHandlerMono::addImage (src/frame_handler_mono.cpp:92 new Frame(cam, img),
 * :190-204 CoarseTracker::run(last_frame, new_frame)). imgs[b] becomes a new device frame (id in new_ids[b], statistics in integral / grad_mean,
 * either may be NULL); jobs[b].cur is ignored and replaced by that frame; jobs[b].exposure_rat < 0 lets the device form
 * cur.integralImage_/ref.integralImage_ like CoarseTracker.cpp:60. Internally the batch is chunk-pipelined: the H2D copies and the host-side
 * flattening of chunk c+1 overlap the pyramid and tracker kernels of chunk c. Results are identical to hso_frame_upload_batch followed by
 * hso_coarse_track_batch on the same chunking. */
int hso_add_frames_track_batch(hso_ctx* ctx, const hso_track_params* prm, int B, const uint8_t* const* imgs, int W, int H, int stride,
                               const hso_track_job* jobs, hso_frame_id* new_ids, float* integral, float* grad_mean, hso_track_result* out);
/* Where the feature arrays of a job (px, f, dist) are flattened to the device layout (features with dist >= 0 only, xyz = f * dist):
 * -1 (default) on the device when the arrays are pinned / registered host memory — they are then DMA-copied as they are, arrays of consecutive
 * jobs that are adjacent in memory in one copy — else by host threads into pinned staging; 0 always on the host; 1 always on the device.
 * Both paths give bit-identical results (one IEEE fp64 multiplication per component). */
int hso_track_set_direct_inputs(hso_ctx* ctx, int mode);
/* Tuning knob of the call above: problems per chunk and number of compute streams the chunks rotate over (0 = default: 111, 3). */
int hso_set_pipeline(hso_ctx* ctx, int chunk, int streams);

/* ---- F3-inner: direct patch matching — replaces the body of bool hso::Matcher::findMatchDirect(const Point&, Frame&,
 * Vector2d&) after the host-side getCloseViewObs/getWarpMatrixAffine (include/hso/matcher.h:153 ; src/matcher.cpp:310-375),
 * i.e. warp::warpAffine (:120-155), feature_alignment::align1D / align2D float overloads
 * (include/hso/feature_alignment.h:47-55,66-73 ; src/feature_alignment.cpp:164-308,464-605), checkNormal (:406-440),
 * checkNCC (:379-404) and the 20 px displacement gate. ------------------------------------------------------------------ */
typedef struct {
  int32_t ref_level, search_level, type /* Feature::FeatureType: 0 corner, 1 edgelet, 2 gradient */, scale_patch;
  double px_ref[2], A_cur_ref[4], grad[2], px_cur[2];
  float exposure_rat;
  float ncc_thresh;      /* checkNCC threshold; 0 = Matcher::findMatchDirect's 0.7 (matcher.cpp:366); findMatchSeed passes 0.8 (:510) */
} hso_align_job;
typedef struct {
  int32_t ok, align_converged;
  double px_cur[2];
  double h_inv;
} hso_align_result;
/* jobs may reference different reference frames: ref_frames[m] is the frame id of job m's reference observation. */
int hso_align_batch(hso_ctx* ctx, hso_frame_id cur, int M, const hso_align_job* jobs, const hso_frame_id* ref_frames,
                    int align_max_iter, hso_align_result* out);

/* ---- N1 (next row): map reprojection + grid selection + direct matching in one call — replaces the data path of
 * Reprojector::reprojectMap after the host has enumerated the points to project (src/reprojector.cpp:88-331): reprojectPoint (:504-529),
 * the per-cell candidate ordering (pointQualityComparator :333-344, list::sort is stable), the three selection passes over grid_.cell_order
 * (:262-303) or reprojectCellAll (:545-615) when fewer than maxFts+50 points project into the image, and for every candidate the whole
 * Matcher::findMatchDirect (src/matcher.cpp:270-375) including warp::getWarpMatrixAffine (:46-72), getBestSearchLevel (:74-85) and
 * cam2world (src/camera.cpp:66-87,169-190,297-300). The host keeps the pointer chasing: Point::getCloseViewObs (src/point.cpp:116-136)
 * picks ref_ftr_, and the side effects on Point counters / Feature creation are applied from the per-candidate flags returned here.
 * Candidates are given in the order the reference calls reprojectPoint. std::random_shuffle's cell order is the caller's (cell_order). ---- */
typedef struct {
  double p_host[3];      /* point->hostFeature_->f * (1.0 / point->idist_)  (reprojector.cpp:508) */
  double px_ref[2];      /* ref_ftr_->px (level 0) */
  double f_ref[3];       /* ref_ftr_->f */
  double grad[2];        /* ref_ftr_->grad (edgelets) */
  double depth_ref;      /* matcher.cpp:298-309: 1/idist_ when ref_ftr_'s frame is the host frame, else |ref frame pos - pt.pos_| */
  int32_t host_pose;     /* index into T_f_w of point->hostFeature_->frame */
  int32_t ref_pose;      /* index into T_f_w of ref_ftr_->frame; < 0: getCloseViewObs returned false (matcher.cpp:276-287) */
  hso_frame_id ref_frame; /* device frame of ref_ftr_->frame */
  int32_t ref_level;     /* ref_ftr_->level */
  int32_t ftr_type;      /* ref_ftr_->type: 0 corner, 1 edgelet, 2 gradient */
  int32_t pt_type;       /* Point::type_: 0 deleted, 1 temporary, 2 candidate, 3 unknown, 4 good (point.h:53) */
  int32_t pt_ftr_type;   /* Point::ftr_type_: 0 gradient, 1 edgelet, 2 corner (point.h:54) */
  int32_t scale_patch;   /* matcher.cpp:317-321: keyframe gap < 4 and |128 a - 128| > 30 */
  float exposure_rat;    /* cur.m_exposure_time / ref.m_exposure_time */
  float pad_;
} hso_reproj_cand;
typedef struct {
  int32_t cell_size, n_cols, n_rows;  /* Reprojector::initializeGrid (reprojector.cpp:58-77) */
  int32_t max_fts;                    /* Config::maxFts() */
  int32_t align_max_iter;             /* Matcher::Options::align_max_iter (10) */
  int32_t pad_;
} hso_reproj_grid;
typedef struct {
  int32_t in_frame;      /* reprojectPoint returned true: the candidate entered cell `cell` */
  int32_t cell;
  int32_t tried;         /* findMatchDirect was called for it; tried && !matched => n_failed_reproj_++ on the host */
  int32_t matched;       /* a Feature(frame, px, search_level) is created */
  int32_t search_level;  /* matcher_.search_level_ */
  int32_t order;         /* position of that Feature in the order of creation (frame->fts_), -1 otherwise */
  int32_t align_ok;      /* diagnostics: what findMatchDirect returns for this candidate, evaluated speculatively for every candidate in a cell */
  int32_t pad_;
  double px[2];          /* reprojected pixel; for tried candidates the pixel findMatchDirect left in Candidate::px */
  double A_cur_ref[4];   /* matcher_.A_cur_ref_ (edgelet gradient update, reprojector.cpp:404-409) */
} hso_reproj_result;
typedef struct {
  int32_t n_in_frame;    /* nFeatures_ */
  int32_t n_matches;     /* n_matches_ */
  int32_t n_trials;      /* n_trials_: every list entry the walk reaches, TYPE_DELETED points included (++n_trials_ precedes the test, reprojector.cpp:361-367,553-559) */
  int32_t used_cell_all; /* 1: the reprojectCellAll branch ran (n_in_frame < max_fts + 50) */
} hso_reproj_summary;
/* T_cur_w: frame->T_f_w_ (3x4 row-major); T_f_w: n_poses keyframe poses (3x4 each). cell_order: a permutation of the n_cols*n_rows cell indices.
 * Limits (HSO_ERR_CAPACITY / HSO_ERR_INVALID beyond them): M <= 16384 candidates, n_cols*n_rows <= 4096 cells; the grid must cover the image. */
int hso_reproject_match(hso_ctx* ctx, hso_frame_id cur, const double T_cur_w[12], int n_poses, const double* T_f_w, int M,
                        const hso_reproj_cand* cands, const hso_reproj_grid* grid, const int32_t* cell_order, hso_reproj_result* out,
                        hso_reproj_summary* summary);
/* Test hook: the selection kernel of the call above alone, on caller-given per-candidate facts — in_frame / cell (what reprojectPoint decided) and
 * align_ok (what findMatchDirect returned). Only cands[i].pt_type / pt_ftr_type are read. Lets the parity tests drive the three passes through
 * every corner (src/reprojector.cpp:262-303) without having to construct images that produce a given match pattern. */
int hso_reproject_select_only(hso_ctx* ctx, int M, const hso_reproj_cand* cands, const int32_t* in_frame, const int32_t* cell, const uint8_t* align_ok,
                              const hso_reproj_grid* grid, const int32_t* cell_order, hso_reproj_result* out, hso_reproj_summary* summary);

/* ---- N3 (next row): depth-filter observation — replaces the per-seed body of DepthFilter::observeDepthRow (src/depth_filter.cpp:580-675) for all
 * seeds of the active frame in one launch: visibility test (:591-607), inverse-depth interval (:616-618), Matcher::doLineStereo
 * (src/matcher.cpp:802-1049 with KLTLimited1D / KLTLimited2D :1296-1606, ZMNCC_F include/hso/vikit/patch_score.h:268-305, warp::createPatch
 * :159-196, depthFromTriangulation :242-255), DepthFilter::computeTau (:539-555) and DepthFilter::updateSeed (:528-537). The host keeps the seed
 * list (std::list<Seed>, seeds_mut_), applies b++ / eplStart / vec_distance / last_update_frame from the results, and calls
 * featureExtractor_->setGridOccpuancy for keyframes. ---- */
typedef struct {
  double px[2], f[3], grad[2];  /* seed.ftr->px, ->f, ->grad */
  hso_frame_id ref_frame;       /* device frame of seed.ftr->frame */
  int32_t ref_pose;             /* index into T_f_w of seed.ftr->frame */
  int32_t level, ftr_type;      /* seed.ftr->level, ->type (0 corner, 1 edgelet, 2 gradient) */
  float mu, sigma2;             /* seed.mu, seed.sigma2 */
  float exposure_rat;           /* active_frame.m_exposure_time / ref_frame.m_exposure_time (matcher.cpp:820) */
  float pad_;
} hso_seed_obs;
typedef struct {
  int32_t is_update;            /* Seed::is_update: the seed is in view of the active frame */
  int32_t is_valid;             /* 0: z_inv_min is NaN => Seed::isValid = false (:619) */
  int32_t res;                  /* doLineStereo: 1 ok, -1 epipolar segment / edgelet angle, -2 triangulation, -3 alignment, -4 score; 0 when !is_update */
  int32_t search_level;         /* matcher.search_level_ -> Seed::last_matched_level */
  int32_t epl_start[2], epl_end[2]; /* Seed::eplStart / eplEnd; (0,0) unless res == 1 (:634-635) */
  float mu, sigma2;             /* after updateSeed (unchanged unless res == 1) */
  double z;                     /* result_depth */
  double px_cur[2];             /* matcher.px_cur_ -> Seed::last_matched_px */
} hso_seed_result;
/* px_error_angle: DepthFilter::px_error_angle_ = atan(px_noise / (2 focal_length)) * 2 (:360-365). align_max_iter: Matcher::Options (10). */
int hso_depth_observe(hso_ctx* ctx, hso_frame_id cur, const double T_cur_w[12], int n_poses, const double* T_f_w, double px_error_angle,
                      int align_max_iter, int S, const hso_seed_obs* seeds, hso_seed_result* out);

/* ---- a13b: seed reprojection — replaces the seed stage of Reprojector::reprojectMap (src/reprojector.cpp:309-328): Reprojector::reprojectorSeed
 * (:531-552: project 1/mu * f of every unconverged-but-narrow seed into the frame, z < 0.001 and the 8-px frame test, grid cell), the per-cell
 * ordering by Seed::sigma2 (seedComparator :346-349, std::list::sort is stable), Reprojector::reprojectorSeeds (:431-503: first seed of a cell
 * that Matcher::findMatchSeed accepts) over grid_.cell_order until n_matches_ reaches Config::maxFts(), and the whole Matcher::findMatchSeed
 * (src/matcher.cpp:442-518: parallax test cos < 0.5, warp matrix at depth 1/mu, search level, warpAffine, exposure scaling whenever
 * |128 a - 128| > 30, align1D / align2D, checkNormal, checkNCC at 0.8, 20-px gate). The host keeps the seed list: it selects the seeds
 * (sqrt(sigma2) < z_range / reproject_seed_thresh && !haveReprojected, :316), decides whether the stage runs at all (n_matches_ < 100) and
 * creates the TYPE_TEMPORARY Points / Features from the results. Seeds use the record of row N3 (hso_seed_obs, above; exposure_rat = frame.m_exposure_time / seed frame's); results use
 * hso_reproj_result (tried = findMatchSeed was called; matched = a Feature is created; order = position among the Features created by this
 * stage). summary->n_matches = n_matches_ after the stage (n_matches_in + new), n_trials = findMatchSeed calls, used_cell_all = 0. ---- */
int hso_reproject_seeds(hso_ctx* ctx, hso_frame_id cur, const double T_cur_w[12], int n_poses, const double* T_f_w, int S, const hso_seed_obs* seeds,
                        const hso_reproj_grid* grid, const int32_t* cell_order, int n_matches_in, hso_reproj_result* out, hso_reproj_summary* summary);

/* ---- F4: pose refinement — replaces void pose_optimizer::optimizeLevenbergMarquardt3rd(double reproj_thresh, size_t n_iter,
 * bool verbose, FramePtr&, double& scale, double& err_init, double& err_final, size_t& num_obs)
 * (include/hso/pose_optimizer.h:61-64 ; src/pose_optimizer.cpp:399-771). ------------------------------------------------- */
typedef struct {
  double T_f_w[12], cov[36];
  double estimated_scale, error_init, error_final;
  uint64_t num_obs;
  float error_in_px;
  int32_t n_trials_total, early_return;
} hso_pose_result;
int hso_pose_optimize(hso_ctx* ctx, double reproj_thresh, int n_iter, int n_fts_total, int F, const double* f, const double* p_host,
                      const int32_t* host_idx, int K, const double* T_host_w, const double* grad, const int8_t* level, const int8_t* ftype,
                      const int8_t* ptype, const double T_f_w_in[12], uint8_t* outlier_out, hso_pose_result* out);
/* Many independent frames in one launch. Arrays are concatenated; offs[B+1] indexes features, hoffs[B+1] indexes host poses. */
int hso_pose_optimize_batch(hso_ctx* ctx, double reproj_thresh, int n_iter, int B, const int32_t* n_fts_total, const int32_t* offs,
                            const double* f, const double* p_host, const int32_t* host_idx, const int32_t* hoffs, const double* T_host_w,
                            const double* grad, const int8_t* level, const int8_t* ftype, const int8_t* ptype, const double* T_f_w_in,
                            uint8_t* outlier_out, hso_pose_result* out);

/* ---- N2 (next row): corner detection — replaces the detector part of FeatureExtractor::fastDetectST(const cv::Mat& imageLevel, int Level)
 * (src/feature_detection.cpp:498-523): fast::fast_corner_detect_9_sse2 + fast_corner_score_9 + fast_nonmax_3x3 (thirdparty/fast), the border
 * filter (:515) and hso::shiTomasiScore (src/vikit/vision.cpp:111-151), on a level of a device-resident frame. Bit-exact corner list in
 * raster order. The host caller keeps the cell bookkeeping and the octree distribution. ------------------------------------------------ */
typedef struct {
  int16_t x, y;        /* level pixel (multiply by 1 << level for KeyPoint coordinates, feature_detection.cpp:521) */
  int32_t score;       /* fast_corner_score_9 */
  float shi_tomasi;    /* hso::shiTomasiScore(imageLevel, x, y) */
} hso_corner;
/* threshold = floor(minThresh_); border = 8 in the reference. count receives the number of corners found (may exceed cap: only cap are written). */
int hso_fast_detect(hso_ctx* ctx, hso_frame_id frame, int level, int threshold, int border, hso_corner* out, int cap, int* count);
/* Levels 0 .. n_levels-1 in one call (what FeatureExtractor::fastDetectMT does on three threads, feature_detection.cpp:498-514): out is
 * [n_levels][cap_per_level], counts[l] the number of corners found on level l. One synchronisation when no level finds more than 4096 corners. */
int hso_fast_detect_levels(hso_ctx* ctx, hso_frame_id frame, int n_levels, int threshold, int border, hso_corner* out, int cap_per_level, int* counts);

/* ---- stage timers, named like the reference's HSO_START_TIMER sites (src/frame_handler_base.cpp:57-66) ------------------- */
/* Accumulated device time in ms and number of calls of stage: 0 "pyramid_creation", 1 "sparse_img_align", 2 "feature_align" (the alignment kernel:
 * hso_align_batch, and nested inside hso_reproject_match / hso_reproject_seeds as in src/reprojector.cpp:259-330), 3 "pose_optimizer",
 * 4 "reproject" (whole hso_reproject_match / hso_reproject_seeds call), 5 "depth_filter_update" (hso_depth_observe; the reference logs it on
 * the mapping thread), 6 "feature_detection" (hso_fast_detect*). "reproject_kfs" / "reproject_candidates" / "local_ba" / "tot_time" time host
 * loops of the caller and stay there. hso_stage_name returns the name, NULL beyond the last stage. */
int hso_stage_time_ms(hso_ctx* ctx, int stage, double* ms, uint64_t* calls);
const char* hso_stage_name(int stage);

#ifdef __cplusplus
}
#endif
#endif
