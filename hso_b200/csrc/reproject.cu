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
  jd.job.exposure_rat = c.exposure_rat; jd.job.ncc_thresh = 0.f;

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

// Exclusive prefix over the block of one value per thread (thread order); *total receives the block sum. s_warp: 32 ints of scratch.
HSO_DEV int block_excl_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // s_warp may still be read from the previous scan
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int before = 0, tot = 0;
  for (int q = 0; q < SEL_THREADS / 32; ++q) {
    const int t = s_warp[q];
    if (q < warp) before += t;
    tot += t;
  }
  *total = tot;
  return before + incl - v;
}

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
  int* ord = base3 + P.n_cells;                                           // grid_.cell_order staged in shared memory
  uint8_t* has1 = reinterpret_cast<uint8_t*>(ord + P.n_cells);            // per-cell summaries
  uint8_t* has2 = has1 + P.n_cells;
  uint8_t* n3 = has2 + P.n_cells;                                         // successes (<= 3) behind the pass-2 position
  uint8_t* n3b = n3 + P.n_cells;                                          // successes (<= 3) behind the pass-1 position (cell_order[0]: the 2nd pass never visits it)
  uint8_t* proc = n3b + P.n_cells;                                        // bit p: pass p+1 visited the cell
  uint8_t* lim3 = proc + P.n_cells;
  uint8_t* flag = lim3 + P.n_cells;                                       // [M] bit 0: findMatchDirect succeeded, bit 1: in frame, bit 2: TYPE_DELETED
  __shared__ int s_warp[32];
  __shared__ int s_n_in, s_trials, s_matches;
  const int tid = threadIdx.x, lane = tid & 31;
  const int M = P.M;

  if (tid == 0) { s_n_in = 0; s_trials = 0; s_matches = 0; }
  __syncthreads();
  {
    int cnt = 0;
    for (int i = tid; i < M; i += SEL_THREADS) {
      const int ok = ar[i].ok != 0, inf = res[i].in_frame != 0, del = cands[i].pt_type == 0;
      res[i].align_ok = ok;
      flag[i] = (uint8_t)((ok && !del ? 1 : 0) | (inf ? 2 : 0) | (del ? 4 : 0));  // a deleted point is erased before findMatchDirect: never a match
      cnt += inf;
    }
    cnt = warp_sum(cnt);
    if (lane == 0 && cnt) atomicAdd(&s_n_in, cnt);
  }
  __syncthreads();
  const int n_in = s_n_in;
  const int maxf = P.max_fts;
  const bool cell_all = n_in < maxf + 50;  // allPixelToDistribute.size() < Config::maxFts()+50 (src/reprojector.cpp:257)

  if (cell_all) {
    // ---- Reprojector::reprojectCellAll (src/reprojector.cpp:545-615): candidates in insertion order until max_fts matches ---------
    const int per = (M + SEL_THREADS - 1) / SEL_THREADS;
    const int i0 = min(M, tid * per), i1 = min(M, i0 + per);
    int local = 0;
    for (int i = i0; i < i1; ++i) local += ((flag[i] & 3) == 3) ? 1 : 0;
    int tot;
    int pre = block_excl_scan(local, s_warp, &tot);  // successes among earlier candidates
    int trials = 0, matches = 0;
    for (int i = i0; i < i1; ++i) {
      const uint8_t f = flag[i];
      if (!(f & 2)) continue;
      const bool reached = pre < max(maxf, 1);  // the loop only returns after a success (n_matches_ >= maxFts is tested there, :611-612)
      if (reached) ++trials;                    // ++n_trials_ comes before the TYPE_DELETED test (:553-559)
      if (f & 4) continue;
      const bool ok = (f & 1) != 0;
      if (reached) {
        res[i].tried = 1;
        res[i].px[0] = ar[i].px_cur[0]; res[i].px[1] = ar[i].px_cur[1];
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
    if (i < M && (flag[i] & 2)) {
      // pointQualityComparator: type desc, then ftr_type desc. TYPE_DELETED (0) points stay in the cell lists (they sort last): the walk
      // counts a trial for each one it reaches before erasing it (src/reprojector.cpp:361-367)
      const int q = cands[i].pt_type * 3 + cands[i].pt_ftr_type;
      k = ((uint32_t)res[i].cell << 20) | ((uint32_t)(15 - q) << 16) | (uint32_t)i;
    }
    keys[i] = k;
  }
  for (int c = tid; c < P.n_cells; c += SEL_THREADS) {
    cs[c] = 0; ce[c] = 0; proc[c] = 0; has1[c] = has2[c] = n3[c] = n3b[c] = lim3[c] = 0;
    ord[c] = cell_order[c];
  }
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
    const int e = ce[c];
    int p = cs[c];
    uint8_t h1 = 0, h2 = 0, k3 = 0, k3b = 0;
    for (; p < e; ++p) if (flag[keys[p] & 0xFFFFu] & 1) { h1 = 1; ++p; break; }
    for (int q = p; q < e && k3b < 3; ++q) if (flag[keys[q] & 0xFFFFu] & 1) ++k3b;  // pass 3 right behind pass 1
    for (; p < e; ++p) if (flag[keys[p] & 0xFFFFu] & 1) { h2 = 1; ++p; break; }
    for (; p < e && k3 < 3; ++p) if (flag[keys[p] & 0xFFFFu] & 1) ++k3;
    has1[c] = h1; has2[c] = h2; n3[c] = k3; n3b[c] = k3b;
  }
  __syncthreads();
  // ---- the three passes over grid_.cell_order (src/reprojector.cpp:262-303) as three block-wide prefix scans over the order positions. A pass
  // stops behind the first position at which n_matches_ reaches maxFts, so position i is visited iff i is the pass's first position or the
  // running count in front of it is still below maxFts; the count a cell's matches start at is that running count. ------------------------------
  const int nc = P.n_cells;
  const int per = (nc + SEL_THREADS - 1) / SEL_THREADS;
  const int p0 = min(nc, tid * per), p1 = min(nc, p0 + per);
  int n1, n2 = 0, n3tot = 0;
  {  // 1st: i = 0 .. n_cells-1, one match per cell
    int local = 0;
    for (int i = p0; i < p1; ++i) local += has1[ord[i]];
    int tot;
    int run = block_excl_scan(local, s_warp, &tot);
    for (int i = p0; i < p1; ++i) {
      const int c = ord[i];
      if (i == 0 || run < maxf) { proc[c] |= 1; base1[c] = run; }
      run += has1[c];
    }
    n1 = min(tot, max(maxf, (int)has1[ord[0]]));  // matches of the visited prefix (position 0 is visited even when maxFts == 0)
  }
  const bool run2 = n1 < maxf, run3_possible = run2;
  if (run2) {  // 2nd: i = n_cells-1 .. 1 — for(size_t i=cells.size()-1; i>0; --i) never visits cell_order[0]
    // thread t owns the mirrored positions so that the scan still runs in thread order: j = n_cells-1-i, j = 0 .. n_cells-2
    const int m = nc - 1;
    const int q0 = min(m, tid * per), q1 = min(m, q0 + per);
    int local = 0;
    for (int j = q0; j < q1; ++j) local += has2[ord[nc - 1 - j]];
    int tot;
    int run = n1 + block_excl_scan(local, s_warp, &tot);
    for (int j = q0; j < q1; ++j) {
      const int c = ord[nc - 1 - j];
      if (j == 0 || run < maxf) { proc[c] |= 2; base2[c] = run; }
      run += has2[c];
    }
    n2 = min(n1 + tot, maxf);  // n1 < maxFts and one match per cell: the count reaches maxFts exactly
  }
  const bool run3 = run3_possible && n2 < maxf;
  if (run3) {  // 3rd: i = 0 .. n_cells-1, up to 3 per cell; n_matches_ is counted inside reprojectCell and saturates at maxFts
    __syncthreads();  // proc bits of pass 2 are complete
    int local = 0;
    for (int i = p0; i < p1; ++i) { const int c = ord[i]; local += (proc[c] & 2) ? n3[c] : n3b[c]; }
    int tot;
    int run = n2 + block_excl_scan(local, s_warp, &tot);
    for (int i = p0; i < p1; ++i) {
      const int c = ord[i];
      const int avail = (proc[c] & 2) ? n3[c] : n3b[c];
      if (i == 0 || run < maxf) { proc[c] |= 4; base3[c] = min(run, maxf); lim3[c] = (uint8_t)max(0, min(3, maxf - run)); }
      run += avail;
    }
    n3tot = min(n2 + tot, maxf);
  }
  const int n_matches = run3 ? n3tot : (run2 ? n2 : n1);
  __syncthreads();
  // ---- mark what the walk tried / matched ------------------------------------------------------------------------------------------------
  int trials = 0;
  for (int c = tid; c < P.n_cells; c += SEL_THREADS) {
    const int e = ce[c];
    int p = cs[c];
    const uint8_t pr = proc[c];
    auto visit = [&](int pos, bool& ok) {
      const int idx = (int)(keys[pos] & 0xFFFFu);
      const uint8_t f = flag[idx];
      ++trials;  // ++n_trials_ comes before the TYPE_DELETED test (:361-367)
      ok = false;
      if (f & 4) return idx;  // erased without a findMatchDirect call
      ok = (f & 1) != 0;
      res[idx].tried = 1;
      res[idx].px[0] = ar[idx].px_cur[0]; res[idx].px[1] = ar[idx].px_cur[1];
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
  if (tid == 0) { summ->n_in_frame = n_in; summ->n_matches = n_matches; summ->n_trials = s_trials; summ->used_cell_all = 0; }
}

// ---- a13b: seed stage ---------------------------------------------------------------------------------------------------------------
// One thread per seed: Reprojector::reprojectorSeed (src/reprojector.cpp:531-552) then the head of Matcher::findMatchSeed
// (src/matcher.cpp:442-470): parallax test, reference in-frame test, T_cur_ref, getWarpMatrixAffine at depth 1/mu, getBestSearchLevel. Emits the
// align job (NCC threshold 0.8; the patch is scaled by the exposure ratio whenever |128 a - 128| > 30, :472-483 — no keyframe-gap condition).
__global__ void __launch_bounds__(128) k_reproject_seed(const ReprojKParams P, const hso_seed_obs* __restrict__ seeds, const uint8_t* const* __restrict__ ref_pyr,
                                                        AlignJobDev* __restrict__ jobs, hso_reproj_result* __restrict__ res) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.M) return;
  const hso_seed_obs s = seeds[i];
  hso_reproj_result r;
  r.in_frame = 0; r.cell = -1; r.tried = 0; r.matched = 0; r.search_level = 0; r.order = -1; r.align_ok = 0; r.pad_ = 0;
  r.px[0] = r.px[1] = 0; r.A_cur_ref[0] = r.A_cur_ref[1] = r.A_cur_ref[2] = r.A_cur_ref[3] = 0;
  AlignJobDev jd;
  jd.ref_pyr = ref_pyr[i];
  jd.job.ref_level = -1;  // skip marker for k_align
  jd.job.search_level = 0; jd.job.type = s.ftr_type;
  jd.job.px_ref[0] = s.px[0]; jd.job.px_ref[1] = s.px[1];
  jd.job.A_cur_ref[0] = jd.job.A_cur_ref[1] = jd.job.A_cur_ref[2] = jd.job.A_cur_ref[3] = 0;
  jd.job.grad[0] = s.grad[0]; jd.job.grad[1] = s.grad[1];
  jd.job.exposure_rat = s.exposure_rat;
  jd.job.scale_patch = fabsf(s.exposure_rat * 128.f - 128.f) > 30.0f ? 1 : 0;  // LIGHT_THRESHOLD (src/matcher.cpp:40,473)
  jd.job.ncc_thresh = 0.8f;                                                    // checkNCC(patch_f_, patchNCC, 0.8) (:510)

  const Se3d Tc = se3_from_rt(P.T_cur_w);
  const Se3d Tr = se3_from_rt(P.T_f_w + 12 * s.ref_pose);
  const Se3d Tri = se3_inverse(Tr);
  const Se3d Tcr = se3_mul(Tc, Tri);  // frame->T_f_w_ * seed.ftr->frame->T_f_w_.inverse()
  const double inv_mu = 1.0 / (double)s.mu;
  // ---- Reprojector::reprojectorSeed (:531-552): pTarget = Tth * (1.0/seed.mu * seed.ftr->f) ------------------------------------------------
  const double hx = inv_mu * s.f[0], hy = inv_mu * s.f[1], hz = inv_mu * s.f[2];
  {
    double X, Y, Z;
    se3_apply(Tcr, hx, hy, hz, X, Y, Z);
    if (!(Z < 0.001)) {
      double pu, pv;
      world2cam_exact(P.cam, X, Y, Z, pu, pv);
      r.px[0] = pu; r.px[1] = pv;
      const int ox = (int)pu, oy = (int)pv;
      if (ox >= 8 && ox < P.cam.width - 8 && oy >= 8 && oy < P.cam.height - 8) {
        r.in_frame = 1;
        r.cell = (int)(pv / (double)P.cell_size) * P.n_cols + (int)(pu / (double)P.cell_size);
      }
    }
  }
  // ---- head of Matcher::findMatchSeed (src/matcher.cpp:442-470) ---------------------------------------------------------------------------
  bool job_ok = r.in_frame != 0;
  if (job_ok) {
    // seed_pos = T_ref^-1 * (1.0/mu * f); ref_dir = (ref frame pos - seed_pos).normalized(); cur_dir likewise; cos < 0.5 => false
    double sx, sy, sz;
    se3_apply(Tri, hx, hy, hz, sx, sy, sz);
    const Se3d Tci = se3_inverse(Tc);
    double rx = Tri.tx - sx, ry = Tri.ty - sy, rz = Tri.tz - sz;  // Frame::pos() = T_f_w_.inverse().translation()
    double cx = Tci.tx - sx, cy = Tci.ty - sy, cz = Tci.tz - sz;
    const double rn = sqrt(rx * rx + ry * ry + rz * rz), cn = sqrt(cx * cx + cy * cy + cz * cz);
    rx /= rn; ry /= rn; rz /= rn; cx /= cn; cy /= cn; cz /= cn;
    const double cos_angle = rx * cx + ry * cy + rz * cz;
    if (cos_angle < 0.5) job_ok = false;
  }
  if (job_ok) {
    const int lv = s.level;
    const int ox = (int)(s.px[0] / (double)(1 << lv)), oy = (int)(s.px[1] / (double)(1 << lv));
    const int boundary = 4 + 2;  // halfpatch_size_ + 2
    job_ok = ox >= boundary && ox < P.cam.width / (1 << lv) - boundary && oy >= boundary && oy < P.cam.height / (1 << lv) - boundary;
  }
  if (job_ok) {
    // warp::getWarpMatrixAffine (src/matcher.cpp:46-72) at depth 1./seed.mu
    const int halfpatch = 5;
    const double xr = s.f[0] * inv_mu, yr = s.f[1] * inv_mu, zr = s.f[2] * inv_mu;
    const int ratio = 1 << s.level;
    double dux, duy, duz, dvx, dvy, dvz;
    cam2world(P.cam, s.px[0] + (double)(halfpatch * ratio), s.px[1], dux, duy, duz);
    cam2world(P.cam, s.px[0], s.px[1] + (double)(halfpatch * ratio), dvx, dvy, dvz);
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
    int sl = 0;
    double D = A[0] * A[3] - A[1] * A[2];
    while (D > 3.0 && sl < P.max_search_level) { sl += 1; D *= 0.25; }
    for (int k = 0; k < 4; ++k) { r.A_cur_ref[k] = A[k]; jd.job.A_cur_ref[k] = A[k]; }
    r.search_level = sl;
    jd.job.ref_level = s.level;
    jd.job.search_level = sl;
  }
  jd.job.px_cur[0] = r.px[0]; jd.job.px_cur[1] = r.px[1];
  jobs[i] = jd;
  res[i] = r;
}

// One CTA: per-cell ordering by (sigma2 ascending, insertion order) — seedComparator with the stable std::list::sort (:346-349,433) — as a bitonic
// sort of 64-bit keys cell | float bits of sigma2 | index; per cell the first seed findMatchSeed accepts (:434-497); the walk over
// grid_.cell_order with its "n_matches_ >= maxFts" break (:320-327) as a block-wide prefix scan.
__global__ void __launch_bounds__(SEL_THREADS) k_seed_select(const SeedSelParams P, const hso_seed_obs* __restrict__ seeds, const hso_align_result* __restrict__ ar,
                                                             const int32_t* __restrict__ cell_order, hso_reproj_result* __restrict__ res,
                                                             hso_reproj_summary* __restrict__ summ) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(sel_smem);  // [n_sort]
  int* cs = reinterpret_cast<int*>(keys + P.n_sort);
  int* ce = cs + P.n_cells;
  int* first_ok = ce + P.n_cells;   // sorted position of the first accepted seed of the cell, -1 if none
  int* base = first_ok + P.n_cells;
  int* ord = base + P.n_cells;
  uint8_t* visited = reinterpret_cast<uint8_t*>(ord + P.n_cells);
  __shared__ int s_warp[32];
  __shared__ int s_n_in, s_trials;
  const int tid = threadIdx.x, lane = tid & 31;
  const int S = P.S;
  if (tid == 0) { s_n_in = 0; s_trials = 0; }
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < P.n_sort; i += SEL_THREADS) {
    unsigned long long k = ~0ull;
    if (i < S) {
      res[i].align_ok = ar[i].ok;
      if (res[i].in_frame) {
        ++cnt;
        // sigma2 > 0: the IEEE bit pattern is monotone; a negative / NaN variance cannot come out of the depth filter
        k = ((unsigned long long)(unsigned)res[i].cell << 48) | ((unsigned long long)__float_as_uint(seeds[i].sigma2) << 16) | (unsigned long long)i;
      }
    }
    keys[i] = k;
  }
  for (int c = tid; c < P.n_cells; c += SEL_THREADS) { cs[c] = 0; ce[c] = 0; first_ok[c] = -1; base[c] = 0; visited[c] = 0; ord[c] = cell_order[c]; }
  cnt = warp_sum(cnt);
  if (lane == 0 && cnt) atomicAdd(&s_n_in, cnt);
  __syncthreads();
  for (int k = 2; k <= P.n_sort; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P.n_sort; i += SEL_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < P.n_sort; i += SEL_THREADS) {
    const unsigned long long k = keys[i];
    if (k == ~0ull) continue;
    const int c = (int)(k >> 48);
    if (i == 0 || (int)(keys[i - 1] >> 48) != c) cs[c] = i;
    if (i + 1 == P.n_sort || keys[i + 1] == ~0ull || (int)(keys[i + 1] >> 48) != c) ce[c] = i + 1;
  }
  __syncthreads();
  for (int c = tid; c < P.n_cells; c += SEL_THREADS)
    for (int p = cs[c]; p < ce[c]; ++p)
      if (ar[keys[p] & 0xFFFFull].ok) { first_ok[c] = p; break; }
  __syncthreads();
  // walk: for i in cell_order: if (reprojectorSeeds(cell)) ++n_matches_; if (n_matches_ >= maxFts) break;
  const int nc = P.n_cells, per = (nc + SEL_THREADS - 1) / SEL_THREADS;
  const int p0 = min(nc, tid * per), p1 = min(nc, p0 + per);
  int local = 0;
  for (int i = p0; i < p1; ++i) local += first_ok[ord[i]] >= 0 ? 1 : 0;
  int tot;
  int run = P.n_matches_in + block_excl_scan(local, s_warp, &tot);
  int trials = 0;
  for (int i = p0; i < p1; ++i) {
    const int c = ord[i];
    const bool vis = (i == 0) || run < P.max_fts;
    if (vis) {
      const int fo = first_ok[c], e = fo >= 0 ? fo + 1 : ce[c];
      for (int p = cs[c]; p < e; ++p) {  // every seed up to and including the first accepted one had findMatchSeed called
        const int idx = (int)(keys[p] & 0xFFFFull);
        res[idx].tried = 1;
        res[idx].px[0] = ar[idx].px_cur[0]; res[idx].px[1] = ar[idx].px_cur[1];
        ++trials;
      }
      if (fo >= 0) {
        const int idx = (int)(keys[fo] & 0xFFFFull);
        res[idx].matched = 1;
        res[idx].order = run - P.n_matches_in;
      }
    }
    run += first_ok[c] >= 0 ? 1 : 0;
  }
  trials = warp_sum(trials);
  if (lane == 0 && trials) atomicAdd(&s_trials, trials);
  __syncthreads();
  if (tid == 0) {
    // matches of the visited prefix: one per cell, so the count reaches maxFts exactly unless the first cell alone overshoots it
    const int first = first_ok[ord[0]] >= 0 ? 1 : 0;
    int n = P.n_matches_in + tot;
    const int cap = max(P.max_fts, P.n_matches_in + first);
    if (n > cap) n = cap;
    summ->n_in_frame = s_n_in; summ->n_matches = n; summ->n_trials = s_trials; summ->used_cell_all = 0;
  }
}

size_t reproj_select_smem(int n_sort, int n_cells) { return sizeof(uint32_t) * n_sort + sizeof(int) * 6 * n_cells + 6 * (size_t)n_cells + (size_t)n_sort + 16; }

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

cudaError_t launch_reproject_seed(const ReprojKParams& p, const hso_seed_obs* seeds_dev, const uint8_t* const* ref_pyr_dev, AlignJobDev* jobs_dev,
                                  hso_reproj_result* res_dev, cudaStream_t stream, uint64_t* launches) {
  k_reproject_seed<<<(p.M + 127) / 128, 128, 0, stream>>>(p, seeds_dev, ref_pyr_dev, jobs_dev, res_dev);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_seed_select(const SeedSelParams& p, const hso_seed_obs* seeds_dev, const hso_align_result* align_dev, const int32_t* cell_order_dev,
                               hso_reproj_result* res_dev, hso_reproj_summary* summ_dev, cudaStream_t stream, uint64_t* launches) {
  const size_t smem = sizeof(unsigned long long) * p.n_sort + sizeof(int) * 5 * p.n_cells + (size_t)p.n_cells + 16;
  cudaError_t e = cudaFuncSetAttribute(k_seed_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_seed_select<<<1, SEL_THREADS, smem, stream>>>(p, seeds_dev, align_dev, cell_order_dev, res_dev, summ_dev);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace hso
