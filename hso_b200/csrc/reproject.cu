// Row N1 on sm_100a: the data path of hso::Reprojector::reprojectMap (src/reprojector.cpp:88-331) for one frame.
//
//   k_reproject      one thread per candidate point: Reprojector::reprojectPoint (:504-529) -> pixel, in-frame test, grid cell; then the
//                    head of Matcher::findMatchDirect (src/matcher.cpp:270-311): reference in-frame test, T_cur_ref, warp::getWarpMatrixAffine
//                    (:46-72, cam2world src/camera.cpp:66-87,169-190,297-300 incl. the cv::undistortPoints iteration), getBestSearchLevel
//                    (:74-85). Emits one align job per candidate in device memory (consumed by k_align, align.cu) — no host round trip.
//   k_align          (align.cu) runs speculatively for EVERY candidate that entered a cell: the reference walks each cell sequentially and
//                    stops at the first success; here all candidates are matched in parallel and the walk is replayed afterwards.
//   k_reproj_select  one CTA: replays the reference's sequential selection exactly — per-cell stable ordering by pointQualityComparator
//                    (:333-344; std::list::sort is a stable merge sort, so the order is (type desc, ftr_type desc, insertion order)) as one
//                    bitonic sort of 32-bit keys in shared memory, per-cell summaries in parallel, the three passes over grid_.cell_order
//                    (:262-303, including the 2nd pass skipping cell_order[0] and the 3-per-cell third pass) as an O(cells) walk by one
//                    thread over the summaries, and reprojectCellAll (:545-615) as a block-wide prefix scan. Outputs which candidates the
//                    reference would have tried / matched and the order in which it would have created the Features.
#include "hso_internal.h"

namespace hso {

__global__ void __launch_bounds__(128) k_reproject(const ReprojKParams P, const hso_reproj_cand* __restrict__ cands,
                                                   const uint8_t* const* __restrict__ ref_pyr, AlignJobDev* __restrict__ jobs,
                                                   hso_reproj_result* __restrict__ res) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.M) return;
  const hso_reproj_cand c = cands[i];
  hso_reproj_result r;
  r.in_frame = 0; r.cell = -1; r.tried = 0; r.matched = 0; r.search_level = 0; r.order = -1; r.align_ok = 0; r.pad_ = 0;
  r.px[0] = r.px[1] = 0; r.A_cur_ref[0] = r.A_cur_ref[1] = r.A_cur_ref[2] = r.A_cur_ref[3] = 0;
  AlignJobDev jd;
  jd.ref_pyr = ref_pyr[i];
  jd.job.ref_level = -1;  // skip marker for k_align
  jd.job.search_level = 0; jd.job.type = c.ftr_type; jd.job.scale_patch = c.scale_patch;
  jd.job.px_ref[0] = c.px_ref[0]; jd.job.px_ref[1] = c.px_ref[1];
  jd.job.A_cur_ref[0] = jd.job.A_cur_ref[1] = jd.job.A_cur_ref[2] = jd.job.A_cur_ref[3] = 0;
  jd.job.grad[0] = c.grad[0]; jd.job.grad[1] = c.grad[1];
  jd.job.px_cur[0] = jd.job.px_cur[1] = 0;
  jd.job.exposure_rat = c.exposure_rat; jd.job.pad_ = 0;

  const Se3d Tc = se3_from_rt(P.T_cur_w);
  // ---- Reprojector::reprojectPoint (src/reprojector.cpp:504-529) ----------------------------------------------------------------
  {
    const Se3d Th = se3_from_rt(P.T_f_w + 12 * c.host_pose);
    const Se3d Tth = se3_mul(Tc, se3_inverse(Th));
    double X, Y, Z;
    se3_apply(Tth, c.p_host[0], c.p_host[1], c.p_host[2], X, Y, Z);
    if (!(Z < 0.00001)) {
      double pu, pv;
      world2cam_exact(P.cam, X, Y, Z, pu, pv);
      r.px[0] = pu; r.px[1] = pv;
      const int ox = (int)pu, oy = (int)pv;  // px.cast<int>()
      if (ox >= 8 && ox < P.cam.width - 8 && oy >= 8 && oy < P.cam.height - 8) {  // isInFrame(px.cast<int>(), 8)
        r.in_frame = 1;
        r.cell = (int)(pv / (double)P.cell_size) * P.n_cols + (int)(pu / (double)P.cell_size);
      }
    }
  }
  // ---- head of Matcher::findMatchDirect (src/matcher.cpp:270-311) -----------------------------------------------------------------
  bool job_ok = r.in_frame && c.pt_type != 0 && c.ref_pose >= 0;
  if (job_ok) {
    const int lv = c.ref_level;
    const int ox = (int)(c.px_ref[0] / (double)(1 << lv)), oy = (int)(c.px_ref[1] / (double)(1 << lv));
    const int boundary = 4 + 2;  // halfpatch_size_ + 2
    job_ok = ox >= boundary && ox < P.cam.width / (1 << lv) - boundary && oy >= boundary && oy < P.cam.height / (1 << lv) - boundary;
  }
  if (job_ok) {
    const Se3d Tr = se3_from_rt(P.T_f_w + 12 * c.ref_pose);
    const Se3d Tcr = se3_mul(Tc, se3_inverse(Tr));
    // warp::getWarpMatrixAffine (src/matcher.cpp:46-72)
    const int halfpatch = 5;
    const double xr = c.f_ref[0] * c.depth_ref, yr = c.f_ref[1] * c.depth_ref, zr = c.f_ref[2] * c.depth_ref;
    const int ratio = 1 << c.ref_level;
    double dux, duy, duz, dvx, dvy, dvz;
    cam2world(P.cam, c.px_ref[0] + (double)(halfpatch * ratio), c.px_ref[1], dux, duy, duz);
    cam2world(P.cam, c.px_ref[0], c.px_ref[1] + (double)(halfpatch * ratio), dvx, dvy, dvz);
    const double sdu = zr / duz, sdv = zr / dvz;
    double ax, ay, az, bx, by, bz, cx, cy, cz;
    se3_apply(Tcr, xr, yr, zr, ax, ay, az);
    se3_apply(Tcr, dux * sdu, duy * sdu, duz * sdu, bx, by, bz);
    se3_apply(Tcr, dvx * sdv, dvy * sdv, dvz * sdv, cx, cy, cz);
    double pcu, pcv, puu, puv, pvu, pvv;
    world2cam_exact(P.cam, ax, ay, az, pcu, pcv);
    world2cam_exact(P.cam, bx, by, bz, puu, puv);
    world2cam_exact(P.cam, cx, cy, cz, pvu, pvv);
    double A[4];
    A[0] = (puu - pcu) / halfpatch; A[2] = (puv - pcv) / halfpatch;
    A[1] = (pvu - pcu) / halfpatch; A[3] = (pvv - pcv) / halfpatch;
    // warp::getBestSearchLevel (src/matcher.cpp:74-85)
    int sl = 0;
    double D = A[0] * A[3] - A[1] * A[2];
    while (D > 3.0 && sl < P.max_search_level) { sl += 1; D *= 0.25; }
    for (int k = 0; k < 4; ++k) { r.A_cur_ref[k] = A[k]; jd.job.A_cur_ref[k] = A[k]; }
    r.search_level = sl;
    jd.job.ref_level = c.ref_level;
    jd.job.search_level = sl;
  }
  jd.job.px_cur[0] = r.px[0]; jd.job.px_cur[1] = r.px[1];  // findMatchDirect leaves px_cur alone when it returns early
  jobs[i] = jd;
  res[i] = r;
}

// ---- selection -------------------------------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 1024;

__global__ void __launch_bounds__(SEL_THREADS) k_reproj_select(const ReprojSelParams P, const hso_reproj_cand* __restrict__ cands,
                                                               const hso_align_result* __restrict__ ar, const int32_t* __restrict__ cell_order,
                                                               hso_reproj_result* __restrict__ res, hso_reproj_summary* __restrict__ summ) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  uint32_t* keys = reinterpret_cast<uint32_t*>(sel_smem);                 // [P.n_sort]
  int* cs = reinterpret_cast<int*>(keys + P.n_sort);                      // [n_cells] segment start in the sorted keys
  int* ce = cs + P.n_cells;                                               // [n_cells] segment end
  int* base1 = ce + P.n_cells;                                            // creation order of the pass-1 / 2 / 3 matches of a cell
  int* base2 = base1 + P.n_cells;
  int* base3 = base2 + P.n_cells;
  uint8_t* has1 = reinterpret_cast<uint8_t*>(base3 + P.n_cells);          // per-cell summaries
  uint8_t* has2 = has1 + P.n_cells;
  uint8_t* n3 = has2 + P.n_cells;
  uint8_t* proc = n3 + P.n_cells;                                         // bit p: pass p+1 visited the cell
  uint8_t* lim3 = proc + P.n_cells;
  __shared__ int s_warp[32];
  __shared__ int s_n_in, s_trials, s_matches;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = P.M;

  if (tid == 0) { s_n_in = 0; s_trials = 0; s_matches = 0; }
  __syncthreads();
  {
    int cnt = 0;
    for (int i = tid; i < M; i += SEL_THREADS) res[i].align_ok = ar[i].ok;
    for (int i = tid; i < M; i += SEL_THREADS) cnt += res[i].in_frame;
    cnt = warp_sum(cnt);
    if (lane == 0 && cnt) atomicAdd(&s_n_in, cnt);
  }
  __syncthreads();
  const int n_in = s_n_in;
  const bool cell_all = n_in < P.max_fts + 50;  // allPixelToDistribute.size() < Config::maxFts()+50 (src/reprojector.cpp:257)

  if (cell_all) {
    // ---- Reprojector::reprojectCellAll (src/reprojector.cpp:545-615): candidates in insertion order until max_fts matches ---------
    const int per = (M + SEL_THREADS - 1) / SEL_THREADS;
    const int i0 = tid * per, i1 = min(M, i0 + per);
    int local = 0;
    for (int i = i0; i < i1; ++i) local += (res[i].in_frame && cands[i].pt_type != 0 && ar[i].ok) ? 1 : 0;
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int q = 0; q < warp; ++q) before += s_warp[q];
    int pre = before + incl - local;  // successes among earlier eligible candidates
    int trials = 0, matches = 0;
    for (int i = i0; i < i1; ++i) {
      const bool eligible = res[i].in_frame && cands[i].pt_type != 0;
      if (!eligible) continue;
      const bool ok = ar[i].ok != 0;
      if (pre < max(P.max_fts, 1)) {  // the loop only returns after a success (n_matches_ >= maxFts is tested there, :611-612)
        res[i].tried = 1;
        res[i].px[0] = ar[i].px_cur[0]; res[i].px[1] = ar[i].px_cur[1];
        ++trials;
        if (ok) { res[i].matched = 1; res[i].order = pre; ++matches; }
      }
      pre += ok ? 1 : 0;
    }
    trials = warp_sum(trials); matches = warp_sum(matches);
    if (lane == 0) { if (trials) atomicAdd(&s_trials, trials); if (matches) atomicAdd(&s_matches, matches); }
    __syncthreads();
    if (tid == 0) { summ->n_in_frame = n_in; summ->n_matches = s_matches; summ->n_trials = s_trials; summ->used_cell_all = 1; }
    return;
  }

  // ---- per-cell ordering: key = cell | inverted quality | insertion index; one bitonic sort ------------------------------------------
  for (int i = tid; i < P.n_sort; i += SEL_THREADS) {
    uint32_t k = 0xFFFFFFFFu;
    if (i < M && res[i].in_frame && cands[i].pt_type != 0) {
      const int q = cands[i].pt_type * 3 + cands[i].pt_ftr_type;  // pointQualityComparator: type desc, then ftr_type desc
      k = ((uint32_t)res[i].cell << 20) | ((uint32_t)(15 - q) << 16) | (uint32_t)i;
    }
    keys[i] = k;
  }
  for (int c = tid; c < P.n_cells; c += SEL_THREADS) { cs[c] = 0; ce[c] = 0; proc[c] = 0; has1[c] = has2[c] = n3[c] = lim3[c] = 0; }
  __syncthreads();
  for (int k = 2; k <= P.n_sort; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P.n_sort; i += SEL_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < P.n_sort; i += SEL_THREADS) {
    const uint32_t k = keys[i];
    if (k == 0xFFFFFFFFu) continue;
    const int c = (int)(k >> 20);
    if (i == 0 || (int)(keys[i - 1] >> 20) != c) cs[c] = i;
    if (i + 1 == P.n_sort || keys[i + 1] == 0xFFFFFFFFu || (int)(keys[i + 1] >> 20) != c) ce[c] = i + 1;
  }
  __syncthreads();
  // ---- per-cell summaries: what each pass would find in this cell -------------------------------------------------------------------
  for (int c = tid; c < P.n_cells; c += SEL_THREADS) {
    const int s = cs[c], e = ce[c];
    int p = s;
    uint8_t h1 = 0, h2 = 0, k3 = 0;
    for (; p < e; ++p) if (ar[keys[p] & 0xFFFFu].ok) { h1 = 1; ++p; break; }
    for (; p < e; ++p) if (ar[keys[p] & 0xFFFFu].ok) { h2 = 1; ++p; break; }
    for (; p < e && k3 < 3; ++p) if (ar[keys[p] & 0xFFFFu].ok) ++k3;
    has1[c] = h1; has2[c] = h2; n3[c] = k3;
  }
  __syncthreads();
  // ---- the three passes over grid_.cell_order (src/reprojector.cpp:262-303), O(cells) on the summaries ------------------------------
  if (tid == 0) {
    int n = 0;
    const int maxf = P.max_fts;
    for (int i = 0; i < P.n_cells; ++i) {  // 1st
      const int c = cell_order[i];
      proc[c] |= 1;
      if (has1[c]) { base1[c] = n; ++n; }
      if (n >= maxf) break;
    }
    if (n < maxf) {  // 2nd: for(size_t i=cells.size()-1; i>0; --i) — never visits cell_order[0]
      for (int i = P.n_cells - 1; i > 0; --i) {
        const int c = cell_order[i];
        proc[c] |= 2;
        if (has2[c]) { base2[c] = n; ++n; }
        if (n >= maxf) break;
      }
    }
    if (n < maxf) {  // 3rd: up to 3 per cell, n_matches_ counted inside reprojectCell
      for (int i = 0; i < P.n_cells; ++i) {
        const int c = cell_order[i];
        proc[c] |= 4;
        const int lim = min(3, maxf - n);
        lim3[c] = (uint8_t)lim;
        base3[c] = n;
        n += min((int)n3[c], lim);
        if (n >= maxf) break;
      }
    }
    s_matches = n;
  }
  __syncthreads();
  // ---- mark what the walk tried / matched ------------------------------------------------------------------------------------------------
  int trials = 0;
  for (int c = tid; c < P.n_cells; c += SEL_THREADS) {
    const int e = ce[c];
    int p = cs[c];
    const uint8_t pr = proc[c];
    auto visit = [&](int pos, bool& ok) {
      const int idx = (int)(keys[pos] & 0xFFFFu);
      ok = ar[idx].ok != 0;
      res[idx].tried = 1;
      res[idx].px[0] = ar[idx].px_cur[0]; res[idx].px[1] = ar[idx].px_cur[1];
      ++trials;
      return idx;
    };
    if (pr & 1) {
      for (; p < e; ++p) { bool ok; const int idx = visit(p, ok); if (ok) { res[idx].matched = 1; res[idx].order = base1[c]; ++p; break; } }
    }
    if (pr & 2) {
      for (; p < e; ++p) { bool ok; const int idx = visit(p, ok); if (ok) { res[idx].matched = 1; res[idx].order = base2[c]; ++p; break; } }
    }
    if (pr & 4) {
      int k = 0;
      const int lim = lim3[c];
      for (; p < e; ++p) {
        bool ok; const int idx = visit(p, ok);
        if (ok) { res[idx].matched = 1; res[idx].order = base3[c] + k; ++k; if (k >= lim) { ++p; break; } }
      }
    }
  }
  trials = warp_sum(trials);
  if (lane == 0 && trials) atomicAdd(&s_trials, trials);
  __syncthreads();
  if (tid == 0) { summ->n_in_frame = n_in; summ->n_matches = s_matches; summ->n_trials = s_trials; summ->used_cell_all = 0; }
}

size_t reproj_select_smem(int n_sort, int n_cells) { return sizeof(uint32_t) * n_sort + sizeof(int) * 5 * n_cells + 5 * (size_t)n_cells + 16; }

cudaError_t launch_reproject(const ReprojKParams& p, const hso_reproj_cand* cands_dev, const uint8_t* const* ref_pyr_dev, AlignJobDev* jobs_dev,
                             hso_reproj_result* res_dev, cudaStream_t stream, uint64_t* launches) {
  k_reproject<<<(p.M + 127) / 128, 128, 0, stream>>>(p, cands_dev, ref_pyr_dev, jobs_dev, res_dev);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_reproj_select(const ReprojSelParams& p, const hso_reproj_cand* cands_dev, const hso_align_result* align_dev,
                                 const int32_t* cell_order_dev, hso_reproj_result* res_dev, hso_reproj_summary* summ_dev, cudaStream_t stream,
                                 uint64_t* launches) {
  const size_t smem = reproj_select_smem(p.n_sort, p.n_cells);
  cudaError_t e = cudaFuncSetAttribute(k_reproj_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_reproj_select<<<1, SEL_THREADS, smem, stream>>>(p, cands_dev, align_dev, cell_order_dev, res_dev, summ_dev);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace hso
