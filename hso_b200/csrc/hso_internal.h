// Internal structures shared between the C-ABI layer (capi.cu) and the kernels. Product code: never includes oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/hso_b200.h"
#include "common.cuh"

namespace hso {

constexpr int kMaxLevels = 5;     // max(Config::nPyrLevels, kltMaxLevel+1) = 5 (src/frame.cpp:92)
constexpr int kMaxPatternN = 25;  // largest pattern (include/hso/CoarseTracker.h:100-109)
constexpr int kMaxCluster = 8;    // portable thread-block-cluster size
constexpr int kPadRows = 3;       // zero rows kept below every pyramid level (gradient taps may touch row == rows)

// Device layout of one frame's pyramid. All frames of a context share the camera, hence the geometry.
// Level l is tightly packed (row stride == w[l], exactly the reference's `stride = img.cols`, src/CoarseTracker.cpp:248),
// starts at the 128-byte aligned offset off[l], and is followed by kPadRows zero rows + 16 zero bytes: the reference's
// forward-mode gradient taps can read row == rows (undefined there; zero here, see DESIGN.md "quirks").
struct PyrGeom {
  int n_levels;
  int w[kMaxLevels], h[kMaxLevels];
  size_t off[kMaxLevels];        // 128-byte aligned
  uint32_t stage_bytes[kMaxLevels];  // bytes a kernel stages in shared memory for level l (image + pad, multiple of 16)
  size_t bytes;                  // total bytes per frame
  int half_path;                 // 1: halfSample chain (W%16==0 && H%16==0, src/frame.cpp:303), 0: cv::resize chain
  int sse_rounding[kMaxLevels];  // level l>=1 is produced with SSE2 double rounding iff w[l-1]%16==0 (src/vikit/vision.cpp:82)
};

// ---- pyramid ------------------------------------------------------------------------------------------------------------
struct PyrJobDev {
  const uint8_t* src;  // raw level-0 image on the device, row stride = src_stride
  uint8_t* pyr;        // destination pyramid buffer (PyrGeom layout)
  int16_t* sobel;      // nullptr, or [level 0..2][gx plane | gy plane], tightly packed w*h each
  double* sums;        // [2 * pyramid_tiles] scratch: per-tile sum |grad|, sum I over the level-0 interior
  float* stats;        // [2] out: Frame::integralImage_, Frame::gradMean_
};
// cv::resize(INTER_LINEAR) coefficient tables for the non-%16 pyramid path, built on the host once per context.
struct ResizeTabDev {
  const int* xofs;     // [dw]
  const short* ialpha; // [2*dw]
  const int* yofs;     // [dh]
  const short* ibeta;  // [2*dh]
  int area_fast;       // exact 2x decimation => (a+b+c+d+2)>>2
};
// counters_dev: [B] zero-initialised uints (tile arrival counters, reset by the kernel). src_aligned16: every job's src pointer
// is 16-byte aligned (enables the TMA row copies). A job whose src equals pyr + off[0] is built in place (no level-0 copy).
cudaError_t launch_pyramid(const PyrGeom& g, const PyrJobDev* jobs_dev, int B, int src_stride, const ResizeTabDev* tabs /*[n_levels]*/,
                           int store_sobel, unsigned int* counters_dev, int src_aligned16, cudaStream_t stream, uint64_t* launches);
int pyramid_tiles(const PyrGeom& g);  // number of level-0 tiles per frame (size of PyrJobDev::sums / 2)

// ---- CoarseTracker ------------------------------------------------------------------------------------------------------
// Per-problem persistent state, lives in device memory across the per-level launches.
struct TrackState {
  Se3d T;        // accepted T_cur_ref
  float a;       // accepted exposure ratio
  int n_iters, n_evals;
  int iters_per_level[8];
  int last_total_terms, last_N;
  unsigned long long patch_evals[8];  // per level: sum over evaluations of patches that produced terms
  int trace_len;
  int pad_;
  unsigned long long cycles[8];       // diagnostics (CTA rank 0, thread 0): launch, setup, control
};

struct TrackJobDev {
  const uint8_t* ref_pyr;
  const uint8_t* cur_pyr;
  int F;                 // features with a valid depth (compacted on the host; see capi.cu)
  int Fpad;              // row length of the SoA scratch arrays (multiple of 32)
  const double* px;      // [2][Fpad]  level-0 pixel of the reference feature
  const double* xyz;     // [3][Fpad]  f * dist
  float* ref_cache;      // [25][Fpad] reference intensities, pattern-major => coalesced for lane == patch
  float* ref_gx;         // [25][Fpad] inverse-compositional reference gradients (nullptr in forward mode)
  float* ref_gy;
  float* absres;         // [25][Fpad] |r| scratch of the robust threshold selection (-1 = not in view)
  uint8_t* vis;          // [Fpad]
  TrackState* state;
  hso_trace* trace;      // [trace_cap] or nullptr
  const float* ref_stats;  // device {integralImage_, gradMean_} of the two frames: the initial exposure ratio is formed on the device
  const float* cur_stats;  // when the caller passes exposure_rat < 0 (a = cur.integralImage_/ref.integralImage_, src/CoarseTracker.cpp:60)
  // direct-input mode: the caller's own arrays as they were copied (px [n_raw][2], f [n_raw][3], dist [n_raw]); k_track_compact builds px / xyz / F
  const double* raw_px;
  const double* raw_f;
  const double* raw_dist;
  int n_raw;
  int pad_;
  // compact layout of the caller (hso_track_job::xyz / px32): xyz [n_raw][3] doubles, px [n_raw][2] floats; every feature is valid
  const double* raw_xyz;
  const float* raw_px32;
};

struct TrackLevelParams {
  int ic, max_level, level, n_iter;
  int trace_cap;
  int w, h;                 // geometry of this level
  size_t level_off;         // byte offset of this level in a pyramid buffer
  int fast;                 // 0: everything in global memory; 1: level image + reference-patch caches of the CTA in shared memory;
                            // 2: (inverse-compositional) current AND reference level in shared memory, no cache
                            // 3: level image in shared memory, reference-patch cache in global memory (L2), streamed through a per-warp ring of 2 buffers
                            // 4: (inverse-compositional) as 3 with ONE buffer per warp (half the shared memory; the other warps cover the fetch)
  int pc;                   // FAST: patch slots per CTA = patches-per-thread * threads
  int absres_smem;          // FAST: the |r| scratch of the threshold selection lives in shared memory ([N][pc] floats)
  int hist_bits;            // radix-select digit width: 11 when the histogram fits next to the caches, else 8
  int cluster;              // CTAs per problem (shared-memory layout depends on it)
  uint32_t img_bytes;       // bytes staged (multiple of 16)
  CamDev cam;
};

size_t track_level_smem_bytes(const TrackLevelParams& p, int threads);
cudaError_t launch_track_level(const TrackLevelParams& p, const TrackJobDev* jobs_dev, int B, int cluster, int threads, cudaStream_t stream,
                               uint64_t* launches);
cudaError_t launch_track_init(const TrackJobDev* jobs_dev, const double* T0 /*[B][12]*/, const float* a0, int B, cudaStream_t stream,
                              uint64_t* launches);
// Device-side flattening of the caller's feature arrays (direct-input mode): keeps the features with dist >= 0 in order, xyz = f * dist.
cudaError_t launch_track_compact(TrackJobDev* jobs_dev, int B, cudaStream_t stream, uint64_t* launches);
cudaError_t launch_track_finish(const TrackJobDev* jobs_dev, hso_track_result* out_dev, int B, cudaStream_t stream, uint64_t* launches);

// ---- align --------------------------------------------------------------------------------------------------------------
struct AlignJobDev {
  hso_align_job job;
  const uint8_t* ref_pyr;
};
cudaError_t launch_align(const PyrGeom& g, const uint8_t* cur_pyr, const int16_t* cur_sobel, const AlignJobDev* jobs_dev, int M, int max_iter,
                         hso_align_result* out_dev, cudaStream_t stream, uint64_t* launches);

// ---- input side (row N4): resize + undistortion remap ---------------------------------------------------------------------------
void build_undistort_maps(const hso_cam& cam, std::vector<short>& map1, std::vector<uint16_t>& map2);  // host, once per camera
cudaError_t launch_remap(const uint8_t* const* srcs_dev, int sw, int sh, int sstride, const short2* map1, const uint16_t* map2,
                         uint8_t* const* dsts_dev, int dw, int dh, int B, cudaStream_t stream, uint64_t* launches);
cudaError_t launch_resize(const uint8_t* const* srcs_dev, int sw, int sh, int sstride, const ResizeTabDev& tab, uint8_t* const* dsts_dev, int dw,
                          int dh, int B, cudaStream_t stream, uint64_t* launches);

// ---- reprojection + grid selection (row N1) -------------------------------------------------------------------------------
struct ReprojKParams {
  CamDev cam;
  double T_cur_w[12];      // frame->T_f_w_
  const double* T_f_w;     // device [n_poses][12]
  int n_poses, M;
  int cell_size, n_cols;
  int max_search_level;    // Config::nPyrLevels() - 1
};
struct ReprojSelParams {
  int M, n_sort /* power of two >= M */, n_cells, max_fts;
};
cudaError_t launch_reproject(const ReprojKParams& p, const hso_reproj_cand* cands_dev, const uint8_t* const* ref_pyr_dev, AlignJobDev* jobs_dev,
                             hso_reproj_result* res_dev, cudaStream_t stream, uint64_t* launches);
// a13b: seed stage of reprojectMap (src/reprojector.cpp:309-328,431-503,531-552 ; src/matcher.cpp:442-518)
cudaError_t launch_reproject_seed(const ReprojKParams& p, const hso_seed_obs* seeds_dev, const uint8_t* const* ref_pyr_dev, AlignJobDev* jobs_dev,
                                  hso_reproj_result* res_dev, cudaStream_t stream, uint64_t* launches);
struct SeedSelParams {
  int S, n_sort /* power of two >= S */, n_cells, max_fts, n_matches_in;
};
cudaError_t launch_seed_select(const SeedSelParams& p, const hso_seed_obs* seeds_dev, const hso_align_result* align_dev, const int32_t* cell_order_dev,
                               hso_reproj_result* res_dev, hso_reproj_summary* summ_dev, cudaStream_t stream, uint64_t* launches);
size_t reproj_select_smem(int n_sort, int n_cells);
cudaError_t launch_reproj_select(const ReprojSelParams& p, const hso_reproj_cand* cands_dev, const hso_align_result* align_dev,
                                 const int32_t* cell_order_dev, hso_reproj_result* res_dev, hso_reproj_summary* summ_dev, cudaStream_t stream,
                                 uint64_t* launches);

// ---- depth filter observation (row N3) ---------------------------------------------------------------------------------------
struct DepthKParams {
  PyrGeom g;
  CamDev cam;
  double T_cur_w[12];      // active_frame_->T_f_w_
  const double* T_f_w;     // device [n_poses][12] keyframe poses
  int S, max_iter, max_search_level;
  double px_error_angle;
  const uint8_t* cur_pyr;
  const int16_t* cur_sobel;  // or nullptr
  size_t sobel_off[3];
};
cudaError_t launch_depth_observe(const DepthKParams& p, const hso_seed_obs* seeds_dev, const uint8_t* const* ref_pyr_dev, hso_seed_result* out_dev,
                                 cudaStream_t stream, uint64_t* launches);

// ---- pose optimiser -----------------------------------------------------------------------------------------------------
struct PoseJobDev {
  int F, K, n_fts_total, pad_;
  const double* f;          // 3F
  const double* p_host;     // 3F
  const int32_t* host_idx;  // F
  const double* T_host_w;   // K x 12
  const double* grad;       // 2F
  const int8_t* level;
  const int8_t* ftype;
  const int8_t* ptype;
  double T_f_w_in[12];
  uint8_t* outlier;         // F
  float* scratch;           // 2F floats for the order statistics
  hso_pose_result* out;
};
struct PoseScratch {  // per problem, device memory
  Se3d* T_host_inv;           // [K]
  double* Tth;                // [K][12]  T_f_w * T_host^-1 for the pose of the pass in flight
  unsigned long long* keys;   // [F] radix-select keys
  int8_t* cls;                // [F] 0 = point-like, 1 = edgelet
};
cudaError_t launch_pose(const PoseJobDev* jobs_dev, const PoseScratch* scratch_dev, int B, double reproj_thresh, int n_iter, double err_mult2,
                        cudaStream_t stream, uint64_t* launches);

// ---- FAST-9 detector (row N2) ---------------------------------------------------------------------------------------------
cudaError_t launch_fast(const uint8_t* level_img, int w, int h, int threshold, int border, int16_t* score_map, uint32_t* rowbuf, int* row_count,
                        hso_corner* out_dev, int cap, int* total_dev, cudaStream_t stream, uint64_t* launches);

}  // namespace hso
