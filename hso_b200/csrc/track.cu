// CoarseTracker on sm_100a: sparse direct image alignment, one thread-block cluster per (ref, cur) problem, the whole
// Levenberg-Marquardt loop of a pyramid level resident on the device (no host round trips between trials).
//
// Replaces hso::CoarseTracker::{precomputeReferencePatches, selectRobustFunctionLevel, computeResiduals, computeGS} and the
// level loop of run() — src/CoarseTracker.cpp:74-195,242-644 — for a batch of independent problems.
//
// Mapping (see DESIGN.md "k_track_level"):
//   * lane == patch. A thread owns patches i = t, t+NT, ... for the whole launch; scratch arrays are [pattern px][patch] (or group-major, below)
//     so that a warp's accesses are coalesced / bank-conflict free.
//   * the current-level image is staged once per launch into shared memory with a TMA bulk copy (cp.async.bulk + mbarrier; SASS UBLKCP).
//     Where the reference-patch cache of precomputeReferencePatches lives is the MODE of the instantiation: 1 = resident in shared memory
//     (split over a cluster when image + cache exceed an SM), 2 = inverse-compositional, both levels resident and the reference samples
//     recomputed per evaluation, 3 / 4 = in global memory (L2 resident), group-major, streamed through a per-warp ring in shared memory by one
//     TMA bulk copy per patch group (two buffers / one buffer per warp) — the footprint no longer depends on the feature count, so every level of the
//     headline configuration runs as ONE CTA per problem and two 256-thread CTAs share an SM at the coarse levels; 0 = image and caches in
//     global memory (level 0 / oversized problems). The host picks the shape per level (capi.cu, track_run_range).
//   * a residual evaluation streams the (2P+4)^2 window of a patch once (rows as aligned words + funnel shift; interpolated image formed once per
//     window position); where only the colour is needed (threshold pass, reference gather, cached inverse-compositional path) the (2P+2)^2 window.
//   * the 7x7 normal equations are not accumulated term by term. Every Jacobian row of a patch has the form
//     J = [-c, gx*A + gy*B] with A,B in R^6 constant over the patch (src/CoarseTracker.cpp:372), so per term only the nine
//     moments  sum w*{gx^2, gx gy, gy^2, c gx, c gy, c^2, r gx, r gy, r c}  are accumulated and the 28+7 entries are
//     expanded once per patch — mathematically identical to computeGS (src/CoarseTracker.cpp:499-525), 4x fewer FMAs.
//   * fp32 inside a patch (as the reference), fp32 tree inside a warp, fp64 across warps / CTAs in a fixed order
//     (run-to-run deterministic); the reference accumulates H in fp32 over all terms (MatrixAccumulator.h:65-140).
//   * the robust thresholds (median / MAD, src/CoarseTracker.cpp:608-630) are exact order statistics: a linear-bin histogram fused with the
//     residual pass, one compaction pass of the chosen bin, a radix select over the compacted list (plain 3-pass radix select on the float bit
//     patterns as the fallback) — the k-th element is order independent, so they match nth_element.
//   * cluster of C CTAs per problem (small batches): partial sums and histograms are exchanged through distributed shared memory;
//     every CTA then runs the (tiny, fp64) damped solve + SE3 update redundantly, so one cluster barrier per trial suffices.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "hso_internal.h"

namespace cg = cooperative_groups;

namespace hso {

// include/hso/CoarseTracker.h:58-120 — first staticPatternNum[idx] entries of every pattern (idx 2 repeats {-1,0}: quirk kept).
__constant__ int8_t c_pat[8][25][2] = {
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
// The same table as compile-time constants: 8 bits per pattern pixel, (dx + 8) | (dy + 8) << 4, eight pixels per 64-bit word. With the
// term loop fully unrolled the offsets fold into immediates (no constant-memory loads in the inner loop).
template <int PIDX> struct PatBits;
template <> struct PatBits<0> { static constexpr unsigned long long w0 = 0x0000000000000088ull, w1 = 0x0000000000000000ull, w2 = 0x0000000000000000ull, w3 = 0x0000000000000000ull; };
template <> struct PatBits<1> { static constexpr unsigned long long w0 = 0x0000009889888778ull, w1 = 0x0000000000000000ull, w2 = 0x0000000000000000ull, w3 = 0x0000000000000000ull; };
template <> struct PatBits<2> { static constexpr unsigned long long w0 = 0x8979988887978777ull, w1 = 0x0000000000000099ull, w2 = 0x0000000000000000ull, w3 = 0x0000000000000000ull; };
template <> struct PatBits<3> { static constexpr unsigned long long w0 = 0x99978a8886797768ull, w1 = 0x00000098898778a8ull, w2 = 0x0000000000000000ull, w3 = 0x0000000000000000ull; };
template <> struct PatBits<4> { static constexpr unsigned long long w0 = 0x99978a8886797768ull, w1 = 0x000000aa6aa666a8ull, w2 = 0x0000000000000000ull, w3 = 0x0000000000000000ull; };
template <> struct PatBits<5> { static constexpr unsigned long long w0 = 0x99978a8886797768ull, w1 = 0x7b9575aa6aa666a8ull, w2 = 0x000000b7b957599bull, w3 = 0x0000000000000000ull; };
template <> struct PatBits<6> { static constexpr unsigned long long w0 = 0x877767a696867666ull, w1 = 0x69a898887868a797ull, w2 = 0x9a8a7a6aa9998979ull, w3 = 0x00000000000000aaull; };
template <> struct PatBits<7> { static constexpr unsigned long long w0 = 0x866646c4a4846444ull, w1 = 0x4ac8a8886848c6a6ull, w2 = 0xac8c6c4ccaaa8a6aull, w3 = 0x00000000000000ccull; };
template <int PIDX>
HSO_DEV constexpr int pat_dx(int n) {
  return (int)(((n < 8 ? PatBits<PIDX>::w0 : n < 16 ? PatBits<PIDX>::w1 : n < 24 ? PatBits<PIDX>::w2 : PatBits<PIDX>::w3) >> ((n & 7) * 8)) & 0xfull) - 8;
}
template <int PIDX>
HSO_DEV constexpr int pat_dy(int n) {
  return (int)(((n < 8 ? PatBits<PIDX>::w0 : n < 16 ? PatBits<PIDX>::w1 : n < 24 ? PatBits<PIDX>::w2 : PatBits<PIDX>::w3) >> ((n & 7) * 8 + 4)) & 0xfull) - 8;
}

constexpr int NRED = 40;       // 28 H + 7 b + E + terms + saturated + patches (+1 pad)
// radix-select digit width: 11 bits (2048 bins; passes 11+11+10) when shared memory allows, else 8 bits (256 bins; 4 passes)

struct TrackCtrl {
  double Rt[12];       // pose used by the evaluation in flight
  Se3d T_acc, T_try;
  float a_acc, a_try, a_eval;
  float lambda, huber, outlier;
  double H[28], b[7];  // accepted system (upper triangle, row-major)
  double step[7];
  double E_old;
  int done, iter, n_err;
  uint32_t sel_prefix, sel_k, sel_n, sel_bin, sel_cnt, list_n;
  uint32_t wsum[32];   // per-warp histogram partial sums of the bucket search
};

struct Smem {
  uint8_t* img;
  float* cache;        // FAST: [N (x3 in IC mode)][pc] reference intensities (+ gradients) of this CTA's patches
  float* absres;       // FAST + absres_smem: [N][pc] |r| of the threshold selection
  uint8_t* vis;        // FAST: [pc]
  double* warp_part;   // [nwarps][NRED]
  double* cta_part;    // [2][NRED]
  double* tot;         // [NRED]
  uint32_t* hist;      // [1 << hist_bits]
  uint32_t* ghist;     // [1 << hist_bits] cluster-wide sum (aliases hist when the cluster is a single CTA)
  TrackCtrl* ctrl;
  uint64_t* mbar;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline size_t smem_layout(uint32_t img_bytes, size_t cache_bytes, size_t abs_bytes, size_t vis_bytes, int nwarps, int hist_bits,
                                              int csize, size_t* o_cache, size_t* o_abs, size_t* o_vis, size_t* o_warp, size_t* o_cta, size_t* o_tot, size_t* o_hist,
                                              size_t* o_ghist, size_t* o_ctrl, size_t* o_mbar) {
  size_t o = align_up(img_bytes, 128);
  *o_cache = o; o += align_up(cache_bytes, 16);
  *o_abs = o; o += align_up(abs_bytes, 16);
  *o_vis = o; o += align_up(vis_bytes, 16);
  *o_warp = o; o += sizeof(double) * nwarps * NRED;
  *o_cta = o; o += sizeof(double) * 2 * NRED;
  *o_tot = o; o += sizeof(double) * NRED;
  *o_hist = o; o += sizeof(uint32_t) << hist_bits;
  *o_ghist = csize > 1 ? o : *o_hist;
  if (csize > 1) o += sizeof(uint32_t) << hist_bits;
  *o_ctrl = o; o += align_up(sizeof(TrackCtrl), 16);
  *o_mbar = o; o += 16;
  return o;
}

// MODE 3 (streamed cache): per warp two buffers of N rows x 32 patches of reference intensities
// behind two mbarriers per warp (padded to 128 B so that the buffers stay 128-byte aligned)
__host__ __device__ inline size_t ring_bar_bytes(int nwarps) { return ((size_t)nwarps * 16 + 127) / 128 * 128; }
__host__ __device__ inline size_t ring_bytes(int N, int nwarps, int depth) { return ring_bar_bytes(nwarps) + (size_t)nwarps * depth * N * 32 * sizeof(float); }

static __host__ __device__ inline int pattern_n(int pidx) { return pidx <= 0 ? 1 : pidx == 1 ? 5 : pidx == 2 ? 9 : pidx <= 4 ? 13 : pidx == 5 ? 21 : 25; }

// Shared memory of one CTA. fast: image + patch caches resident; pc = patch slots per CTA (patches-per-thread * threads).
size_t track_level_smem_bytes(const TrackLevelParams& p, int threads) {
  size_t a, b, c, d, e, f, g, h, i, j;
  const int N = pattern_n(p.max_level - p.level + 2);
  const size_t absb = (p.fast && p.absres_smem) ? (size_t)N * p.pc * sizeof(float) : 0;
  const size_t cache = p.fast == 1 ? (size_t)N * (p.ic ? 3 : 1) * p.pc * sizeof(float) : p.fast >= 3 ? ring_bytes(N * (p.ic ? 3 : 1), threads / 32, p.fast == 4 ? 1 : 2) : 0;
  const uint32_t img = p.fast == 2 ? 2 * (uint32_t)align_up(p.img_bytes, 128) : (p.fast ? p.img_bytes : 0);
  return smem_layout(img, cache, absb, p.fast ? (size_t)p.pc : 0, threads / 32, p.hist_bits, p.cluster, &a, &j, &b, &c, &d, &e, &f, &g, &h, &i);
}

// ---- unaligned 4-byte window from a byte image: two aligned words + funnel shift -----------------------------------------
template <bool STAGE>
HSO_DEV uint32_t ld4(const uint8_t* img, int byte_addr) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(img) + (byte_addr >> 2);
  uint32_t lo, hi;
  if (STAGE) { lo = w[0]; hi = w[1]; }
  else { lo = __ldg(w); hi = __ldg(w + 1); }
  return __funnelshift_r(lo, hi, (byte_addr & 3) << 3);
}
// byte k of w as float, exact, without the conversion (XU) pipe: PRMT builds the bit pattern of 2^23 + byte, FADD removes 2^23.
// (I2F.U8 is quarter rate; the first profile showed it as the top stall reason of the term loop.)
template <int K>
HSO_DEV float byte_to_float(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | K)) - 8388608.0f; }
HSO_DEV float byte_to_float_k(uint32_t w, int k) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | (uint32_t)k)) - 8388608.0f; }
HSO_DEV float b0(uint32_t w) { return byte_to_float<0>(w); }
HSO_DEV float b1(uint32_t w) { return byte_to_float<1>(w); }
HSO_DEV float b2(uint32_t w) { return byte_to_float<2>(w); }
HSO_DEV float b3(uint32_t w) { return byte_to_float<3>(w); }

// H and E are sums of (mostly) same-sign terms: fp32 is enough. b = -sum J r w is a gradient that cancels towards zero at the
// optimum, so it is kept in fp64 end to end (the reference also builds b in double, src/CoarseTracker.cpp:520).
struct Acc {
  float h[28], E;
  double b[7];
  int terms, sat, patches;
};

HSO_DEV void acc_zero(Acc& a) {
#pragma unroll
  for (int i = 0; i < 28; ++i) a.h[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 7; ++i) a.b[i] = 0.0;
  a.E = 0.f; a.terms = 0; a.sat = 0; a.patches = 0;
}

// Projection head shared by computeResiduals / selectRobustFunctionLevel (src/CoarseTracker.cpp:290-323, :557-583).
struct Proj {
  bool ok;
  int base;  // byte index of (u_i, v_i) in the level image
  float wtl, wtr, wbl, wbr;
  double x, y, z;
};

HSO_DEV Proj project_patch(const double* Rt, const CamDev& cam, double X, double Y, double Z, float scale, int border, int w, int h) {
  Proj p;
  p.ok = false;
  p.x = Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[3];
  p.y = Rt[4] * X + Rt[5] * Y + Rt[6] * Z + Rt[7];
  p.z = Rt[8] * X + Rt[9] * Y + Rt[10] * Z + Rt[11];
  if (p.z < 0) return p;
  double pu, pv;
  world2cam(cam, p.x, p.y, p.z, pu, pv);
  const float u = (float)pu * scale, v = (float)pv * scale;
  const float uf = floorf(u), vf = floorf(v);
  // floorf + saturating conversion: NaN/inf projections (z == 0) fail the bounds test instead of being undefined
  const int ui = __float2int_rd(u), vi = __float2int_rd(v);
  if (!(ui >= border && vi >= border && ui < w - border && vi < h - border)) return p;  // overflow-safe form of :310
  const float su = u - uf, sv = v - vf;
  p.wtl = (float)((1.0 - su) * (1.0 - sv));
  p.wtr = (float)(su * (1.0 - sv));
  p.wbl = (float)((1.0 - su) * sv);
  p.wbr = su * sv;
  p.base = vi * w + ui;
  p.ok = true;
  return p;
}

// Frame::jacobian_xyz2uv (include/hso/frame.h:192-212), rows pre-multiplied by fx*scale / fy*scale (CoarseTracker.cpp:372).
HSO_DEV void patch_jacobian(double x, double y, double z, float fxl, float fyl, float* A, float* B) {
  const float zi = (float)(1.0 / z);
  const float xf = (float)x, yf = (float)y;
  const float zi2 = zi * zi;
  const float j02 = xf * zi2, j12 = yf * zi2;
  const float j03 = yf * j02;
  A[0] = -zi * fxl; A[1] = 0.f; A[2] = j02 * fxl; A[3] = j03 * fxl; A[4] = -(1.0f + xf * j02) * fxl; A[5] = yf * zi * fxl;
  B[0] = 0.f; B[1] = -zi * fyl; B[2] = j12 * fyl; B[3] = (1.0f + yf * j12) * fyl; B[4] = -j03 * fyl; B[5] = -xf * zi * fyl;
}

struct Moments { float xx, xy, yy, cx, cy, cc, rx, ry, rc; };  // fp32 over the <= 25 terms of one patch

HSO_DEV void expand_patch(Acc& a, const Moments& m, const float* A, const float* B) {
  a.h[0] += m.cc;
  float P[6], Q[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    a.h[1 + k] -= m.cx * A[k] + m.cy * B[k];
    P[k] = m.xx * A[k] + m.xy * B[k];
    Q[k] = m.xy * A[k] + m.yy * B[k];
    a.b[1 + k] -= (double)m.rx * (double)A[k] + (double)m.ry * (double)B[k];
  }
  a.b[0] += (double)m.rc;
  int idx = 7;
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int k = j; k < 6; ++k) a.h[idx++] += A[j] * P[k] + B[j] * Q[k];
}

// ---- window evaluation -------------------------------------------------------------------------------------------------------
// All N pattern pixels of a patch and the +-1 taps of their central-difference gradients live in one (2P+4)^2 pixel window. Instead of
// fetching 12 taps per pattern pixel (8 LDS + 12 conversions + 20 FMA each), the window is streamed row by row ONCE per patch:
// rows are loaded as aligned words + funnel shift, converted to float once, and the bilinearly interpolated image
//   Ib(x,y) = wtl p(x,y) + wtr p(x+1,y) + wbl p(x,y+1) + wbr p(x+1,y+1)
// is formed once per window position (same expression, same operand order as the reference, src/CoarseTracker.cpp:339-342); then
//   colour = Ib(x,y),  dx = 0.5 (Ib(x+1,y) - Ib(x-1,y)),  dy = 0.5 (Ib(x,y+1) - Ib(x,y-1))      (:368-371, identical expression trees).
// The loops are fully unrolled with a compile-time pattern, so positions no pattern pixel needs are dead code and disappear.
template <int PIDX>
struct PG {
  static constexpr int N = (PIDX == 2) ? 9 : (PIDX == 3 || PIDX == 4) ? 13 : (PIDX == 5) ? 21 : 25;
  static constexpr int P = (PIDX == 5) ? 3 : (PIDX == 7) ? 4 : (PIDX <= 2) ? 1 : 2;
  static constexpr int WIN = 2 * P + 4;        // window edge in pixels, origin at (u_i - P - 1, v_i - P - 1)
  static constexpr int NW = (WIN + 6) / 4;     // words loaded per row (any byte alignment)
  static constexpr int NA = (WIN + 3) / 4;     // aligned words kept per row
};

template <int PIDX, bool SM>
HSO_DEV void win_row(const uint8_t* img, int byte_addr, uint32_t* out) {
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(img) + (byte_addr >> 2);
  const int sh = (byte_addr & 3) << 3;
  uint32_t wv[PG<PIDX>::NW];
#pragma unroll
  for (int j = 0; j < PG<PIDX>::NW; ++j) wv[j] = SM ? wp[j] : __ldg(wp + j);
#pragma unroll
  for (int j = 0; j < PG<PIDX>::NA; ++j) out[j] = __funnelshift_r(wv[j], wv[j + 1], sh);
}

struct TermCtx { float a, huber, cutoff, max_energy; };

// one residual term: Huber weight, energy, the nine moments (src/CoarseTracker.cpp:345-404 fused with computeGS :499-525). TOP: the level is
// m_max_level (no outlier cut-off, energy hw r^2), a launch constant turned into a template parameter.
template <bool TOP>
HSO_DEV void accumulate_term(const TermCtx& t, float c, float color, float gx, float gy, Moments& m, float& Ep, int& sat) {
  // residual = cur_color - (exposure_rat * ref + b) with b == 0 (:346): one fused multiply-add, as the reference's own build contracts it
  const float r = fmaf(-t.a, c, color);
  const float ar = fabsf(r);
  // Huber weight hw = huber / |r| (:348) through the approximate reciprocal (1 ulp): hw only scales terms
  float rcp;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(ar));
  const float hw = ar < t.huber ? 1.f : t.huber * rcp;
  // branch-free form of :350-361: a saturated term adds max_energy and contributes nothing to H, b
  const bool saturated = !TOP && ar > t.cutoff;
  const float hr2 = hw * r * r;
  const float e_in = TOP ? hr2 : hr2 * (2.f - hw);
  Ep += saturated ? t.max_energy : e_in;
  if (!TOP) sat += saturated ? 1 : 0;
  const float w = saturated ? 0.f : hw;
  const float wgx = w * gx, wgy = w * gy, wc = w * c;
  m.xx += wgx * gx; m.xy += wgx * gy; m.yy += wgy * gy;
  m.cx += wc * gx;  m.cy += wc * gy;  m.cc += wc * c;
  m.rx += wgx * r;  m.ry += wgy * r;  m.rc += wc * r;
}

// Streams the window of one patch of image `img` (bilinear weights w*, integer base pixel `base`) and calls
// f(n, Ib, 2 dx, 2 dy) for every pattern pixel n — the central differences WITHOUT their factor 0.5: a power-of-two factor commutes with every
// rounding of the moment sums, so the caller applies 0.25 / 0.5 once per patch (moments_unscale) instead of two multiplications per term
// and gets bit-identical moments. Used for the current image in forward mode and for the reference image in the
// inverse-compositional dual-image mode.
template <int PIDX, bool SM, class F>
HSO_DEV void window_samples(const uint8_t* img, int base, int w, float wtl, float wtr, float wbl, float wbr, F&& f) {
  constexpr int P = PG<PIDX>::P, WIN = PG<PIDX>::WIN, N = PG<PIDX>::N;
  const int a0 = base - (P + 1) * w - (P + 1);
  float pxPrev[WIN], pxCur[WIN];
  float IbA[WIN - 1], IbB[WIN - 1], IbC[WIN - 1];  // interpolated rows cy-1, cy, cy+1
#pragma unroll
  for (int r = 0; r < WIN; ++r) {
    uint32_t row[PG<PIDX>::NA];
    win_row<PIDX, SM>(img, a0 + r * w, row);
#pragma unroll
    for (int c = 0; c < WIN; ++c) pxCur[c] = byte_to_float_k(row[c >> 2], c & 3);
    if (r >= 1) {
#pragma unroll
      for (int c = 0; c < WIN - 1; ++c) {
        IbA[c] = IbB[c];
        IbB[c] = IbC[c];
        IbC[c] = wtl * pxPrev[c] + wtr * pxPrev[c + 1] + wbl * pxCur[c] + wbr * pxCur[c + 1];
      }
    }
    if (r >= 3) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        if (pat_dy<PIDX>(n) + P + 1 == r - 2) {
          const int cx = pat_dx<PIDX>(n) + P + 1;
          f(n, IbB[cx], IbB[cx + 1] - IbB[cx - 1], IbC[cx] - IbA[cx]);  // 2 dx, 2 dy: the caller rescales the patch's moments (exact)
        }
      }
    }
#pragma unroll
    for (int c = 0; c < WIN; ++c) pxPrev[c] = pxCur[c];
  }
}

// Colour-only variant: the bilinearly interpolated image at the N pattern pixels, f(n, Ib), from the (2P+2)^2 window — rows loaded once as aligned
// words instead of two unaligned 4-byte fetches per pattern pixel (4 shared-memory loads per term at ~3.4 bank conflicts each made the
// inverse-compositional cached path and the threshold pass LSU bound). Same expression, same operand order as ld4 + the four products (:339-342).
template <int PIDX, bool SM, class F>
HSO_DEV void window_colors(const uint8_t* img, int base, int w, float wtl, float wtr, float wbl, float wbr, F&& f) {
  constexpr int P = PG<PIDX>::P, WIN = 2 * P + 2, N = PG<PIDX>::N;
  constexpr int NW = (WIN + 6) / 4, NA = (WIN + 3) / 4;
  const int a0 = base - P * w - P;
  float pxPrev[WIN], pxCur[WIN];
#pragma unroll
  for (int r = 0; r < WIN; ++r) {
    const int byte_addr = a0 + r * w;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(img) + (byte_addr >> 2);
    const int sh = (byte_addr & 3) << 3;
    uint32_t wv[NW], row[NA];
#pragma unroll
    for (int j = 0; j < NW; ++j) wv[j] = SM ? wp[j] : __ldg(wp + j);
#pragma unroll
    for (int j = 0; j < NA; ++j) row[j] = __funnelshift_r(wv[j], wv[j + 1], sh);
#pragma unroll
    for (int c = 0; c < WIN; ++c) pxCur[c] = byte_to_float_k(row[c >> 2], c & 3);
    if (r >= 1) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        if (pat_dy<PIDX>(n) + P == r - 1) {
          const int cx = pat_dx<PIDX>(n) + P;
          f(n, wtl * pxPrev[cx] + wtr * pxPrev[cx + 1] + wbl * pxCur[cx] + wbr * pxCur[cx + 1]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < WIN; ++c) pxPrev[c] = pxCur[c];
  }
}

// moments accumulated with gradients 2 dx, 2 dy -> moments of dx, dy (exact: powers of two)
HSO_DEV void moments_unscale(Moments& m) {
  m.xx *= 0.25f; m.xy *= 0.25f; m.yy *= 0.25f;
  m.cx *= 0.5f; m.cy *= 0.5f; m.rx *= 0.5f; m.ry *= 0.5f;
}

struct LevelCtx {
  const uint8_t* cur;  // shared-memory copy (FAST) or global level image
  int w, h, border;
  float scale, fxl, fyl;
  bool top;
};

// Where the per-patch caches of the calling thread live. FAST: shared memory, slot = k * blockDim + tid (conflict free);
// SLOW: global scratch, slot = patch index (coalesced). Layout [pattern px][stride].
struct PatchStore {
  const uint8_t* ref;  // MODE 2 (dual image): the reference level in shared memory; intensities/gradients are recomputed per evaluation
  float* cache;    // reference intensities
  float* ring;     // MODE 3: this warp's two staging buffers [2][N][32] for the cache blocks streamed from global memory (L2)
  uint64_t* ring_bar;  // MODE 3: this warp's two mbarriers (one per buffer)
  float* gx;       // inverse-compositional reference gradients
  float* gy;
  uint8_t* vis;
  int stride;
};

template <bool FAST>
HSO_DEV int slot_of(int i, int k) { return FAST ? (k * (int)blockDim.x + (int)threadIdx.x) : i; }

// Reference-side geometry of one patch (src/CoarseTracker.cpp:437-467): integer base pixel and bilinear weights, where
// w_br = 1 - (tl + tr + bl) (quirk: differs from the current-image weights, :467 vs :323).
struct RefPatch { bool in; int base; float wtl, wtr, wbl, wbr; };
HSO_DEV RefPatch ref_patch(double pxu, double pxv, float scale, int border, int w, int h) {
  RefPatch r;
  const float u = (float)(pxu * (double)scale), v = (float)(pxv * (double)scale);
  const float uf = floorf(u), vf = floorf(v);
  const int ui = __float2int_rd(u), vi = __float2int_rd(v);
  r.in = ui >= border && vi >= border && ui < w - border && vi < h - border;  // :441
  const float su = u - uf, sv = v - vf;
  r.wtl = (float)((1.0 - su) * (1.0 - sv));
  r.wtr = (float)(su * (1.0 - sv));
  r.wbl = (float)((1.0 - su) * sv);
  r.wbr = (float)(1.0 - (double)(r.wtl + r.wtr + r.wbl));
  r.base = vi * w + ui;
  return r;
}
// reference intensity of one pattern pixel (:482-485)
template <bool SM>
HSO_DEV float ref_intensity(const uint8_t* img, const RefPatch& r, int addr, int w) {
  const uint32_t r0 = ld4<SM>(img, addr);
  const uint32_t r1 = ld4<SM>(img, addr + w);
  return r.wtl * b0(r0) + r.wtr * b1(r0) + r.wbl * b0(r1) + r.wbr * b1(r1);
}

// One residual evaluation over the calling thread's patches: computeResiduals + computeGS fused
// (src/CoarseTracker.cpp:242-414, :499-525).
// MODE 3: the reference-intensity cache of a problem stays in global memory (L2 resident: it is written once per level and re-read by every
// evaluation) and is streamed through a small per-warp ring in shared memory: the cache is stored GROUP-major — the N x 32 floats of 32 consecutive
// patches are one contiguous block — so that ONE TMA bulk copy (cp.async.bulk + mbarrier, issued by lane 0) fetches the warp's NEXT patch group
// while the current group is evaluated. That frees the 108-258 KB the resident cache takes, so that a level whose image + cache exceed one SM
// (level 1 at 640x480 with 3000 patches) still runs as ONE CTA per problem: no cluster barriers, no DSMEM exchange, one control step and one
// staged image per problem instead of two.
template <int N>
HSO_DEV int ring_index(int i, int n) { return (i >> 5) * (N * 32) + n * 32 + (i & 31); }  // element (patch i, pattern pixel n) of the group-major cache
// g: running number of the group within the launch; i0: first patch of the group (multiple of 32). Depth 2: buffer g & 1, mbarrier phase
// (g >> 1) & 1, group g + 1 is issued before group g is waited for. Depth 1 (a ring of 3 N rows for 16 warps does not fit twice beside a 77 KB
// image): one buffer, phase g & 1, group g + 1 is issued once every lane is done with group g — the other warps of the CTA cover the fetch.
template <int N, int DEPTH>
HSO_DEV void ring_issue(const PatchStore& ps, int i0, uint32_t g) {
  if ((threadIdx.x & 31) == 0) {
    const uint32_t buf = DEPTH == 2 ? (g & 1) : 0;
    uint64_t* bar = ps.ring_bar + buf;
    fence_proxy_async();  // the buffer was last read through the generic proxy (every lane is past it: __syncwarp at the end of that iteration)
    mbar_expect_tx(bar, N * 128);
    tma_bulk_g2s(ps.ring + buf * (N * 32), ps.cache + (size_t)(i0 >> 5) * (N * 32), N * 128, bar);
  }
}
template <int N, int DEPTH>
HSO_DEV const float* ring_wait(const PatchStore& ps, uint32_t g) {
  const uint32_t buf = DEPTH == 2 ? (g & 1) : 0;
  mbar_wait(ps.ring_bar + buf, DEPTH == 2 ? (g >> 1) & 1 : g & 1);
  return ps.ring + buf * (N * 32) + (threadIdx.x & 31);
}

template <int PIDX, bool IC, int MODE, bool TOP>
HSO_DEV void eval_patches(const LevelCtx& L, const TrackJobDev& job, const PatchStore& ps, const CamDev& cam, const double* Rt, float a, float huber,
                          float cutoff, int t0, int nt, Acc& acc, uint32_t& ring_g) {
  constexpr int N = (PIDX == 2) ? 9 : (PIDX == 3 || PIDX == 4) ? 13 : (PIDX == 5) ? 21 : 25;
  constexpr bool FAST = MODE != 0, DUAL = MODE == 2, STREAM = MODE >= 3;
  constexpr int NR = IC ? 3 * N : N;  // rows of a streamed cache group: intensities (+ the two gradient planes, inverse-compositional)
  constexpr int RD = MODE == 4 ? 1 : 2;  // ring buffers per warp
  const int Fp = job.Fpad, S = ps.stride;
  const int lane = threadIdx.x & 31;
  uint32_t g = ring_g;
  if (STREAM && t0 - lane < job.F) ring_issue<NR, RD>(ps, t0 - lane, g);
  // max_energy = 2*huber*cutoff - huber^2, evaluated in double like the reference (cutoff_error is a double there)
  const float max_energy = (float)(2.0 * (double)huber * (double)cutoff - (double)(huber * huber));
  // geometry of the next patch is fetched while the current one is processed (the only global loads of the FAST path)
  int i = t0, k = 0;
  bool have = i < job.F && ps.vis[slot_of<FAST>(i, k)] != 0;
  double X = 0, Y = 0, Z = 1, PU = 0, PV = 0;
  if (have) {
    X = job.xyz[i]; Y = job.xyz[Fp + i]; Z = job.xyz[2 * Fp + i];
    if (DUAL) { PU = job.px[i]; PV = job.px[Fp + i]; }
  }
  // (STREAM: the trip count is warp uniform — the ring is filled and drained by the whole warp)
  while ((STREAM ? i - lane : i) < job.F) {
    const int in = i + nt, kn = k + 1;
    const float* cbuf = nullptr;
    if (STREAM) {
      // the next group goes into the buffer the previous group was read from (every lane is past it: __syncwarp at the end of the iteration)
      if (RD == 2 && in - lane < job.F) ring_issue<NR, RD>(ps, in - lane, g + 1);
      cbuf = ring_wait<NR, RD>(ps, g);
    }
    const bool have_n = in < job.F && ps.vis[slot_of<FAST>(in, kn)] != 0;
    double Xn = 0, Yn = 0, Zn = 1, PUn = 0, PVn = 0;
    if (have_n) {
      Xn = job.xyz[in]; Yn = job.xyz[Fp + in]; Zn = job.xyz[2 * Fp + in];
      if (DUAL) { PUn = job.px[in]; PVn = job.px[Fp + in]; }
    }
    if (have) {
      const Proj p = project_patch(Rt, cam, X, Y, Z, L.scale, L.border, L.w, L.h);
      if (p.ok) {
        const int sl = slot_of<FAST>(i, k);
        Moments m = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        float Ep = 0.f;
        int sat = 0;
        TermCtx tc;
        tc.a = a; tc.huber = huber; tc.cutoff = cutoff; tc.max_energy = max_energy;
        if (!IC) {
          // forward mode: colour and gradients of the CURRENT image from one streamed window
          window_samples<PIDX, FAST>(L.cur, p.base, L.w, p.wtl, p.wtr, p.wbl, p.wbr, [&](int n, float color, float gx, float gy) {
            accumulate_term<TOP>(tc, STREAM ? cbuf[n * 32] : ps.cache[n * S + sl], color, gx, gy, m, Ep, sat);
          });
        } else if (DUAL) {
          // inverse-compositional, dual image: intensity and gradients of the REFERENCE image from one streamed window,
          // the current colour from its four taps
          const RefPatch rp = ref_patch(PU, PV, L.scale, L.border, L.w, L.h);
          window_samples<PIDX, true>(ps.ref, rp.base, L.w, rp.wtl, rp.wtr, rp.wbl, rp.wbr, [&](int n, float c, float gx, float gy) {
            const int addr = p.base + pat_dy<PIDX>(n) * L.w + pat_dx<PIDX>(n);
            const uint32_t r0 = ld4<true>(L.cur, addr);
            const uint32_t r1 = ld4<true>(L.cur, addr + L.w);
            const float color = p.wtl * b0(r0) + p.wtr * b1(r0) + p.wbl * b0(r1) + p.wbr * b1(r1);
            accumulate_term<TOP>(tc, c, color, gx, gy, m, Ep, sat);
          });
        } else {
          // inverse-compositional, cached reference intensities and gradients: only the current colour is sampled
          window_colors<PIDX, FAST>(L.cur, p.base, L.w, p.wtl, p.wtr, p.wbl, p.wbr, [&](int n, float color) {
            if (STREAM) accumulate_term<TOP>(tc, cbuf[n * 32], color, cbuf[(N + n) * 32], cbuf[(2 * N + n) * 32], m, Ep, sat);
            else accumulate_term<TOP>(tc, ps.cache[n * S + sl], color, ps.gx[n * S + sl], ps.gy[n * S + sl], m, Ep, sat);
          });
        }
        if (!IC || DUAL) moments_unscale(m);  // the streamed window delivers 2 dx, 2 dy
        // Jacobian rows of the patch after the term loop (keeps 12 registers free while the window is live)
        float A[6], B[6];
        if (!IC) {
          patch_jacobian(p.x, p.y, p.z, L.fxl, L.fyl, A, B);
        } else {
          patch_jacobian(X, Y, Z, L.fxl, L.fyl, A, B);
#pragma unroll
          for (int q = 0; q < 6; ++q) { A[q] *= a; B[q] *= a; }  // m_jacobian_cache_true = exposure_rat * raw (:244-245)
        }
        expand_patch(acc, m, A, B);
        acc.E += Ep;
        acc.terms += N;
        acc.sat += sat;
        acc.patches += 1;
      }
    }
    i = in; k = kn; have = have_n; X = Xn; Y = Yn; Z = Zn; PU = PUn; PV = PVn;
    if (STREAM) {
      __syncwarp();
      if (RD != 2 && i - lane < job.F) ring_issue<NR, RD>(ps, i - lane, g + 1);  // (i is already the next group's patch)
      ++g;
    }
  }
  ring_g = g;
}

// CTA + cluster reduction of the per-thread partial sums into s.tot[0..NRED) (identical in every CTA of the cluster).
HSO_DEV void reduce_acc(const Acc& acc, const Smem& s, int slot, int csize, int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* wp = s.warp_part + warp * NRED;
  // Transposed butterfly: 32 fp32 quantities over 32 lanes in 16+8+4+2+1 = 31 shuffles (instead of 32 x 5): at every step a lane
  // keeps the half of the values whose index bit matches its lane bit and sends the other half; lane l ends with the total of value l.
  float v[32];
#pragma unroll
  for (int k = 0; k < 28; ++k) v[k] = acc.h[k];
  v[28] = acc.E; v[29] = (float)acc.terms; v[30] = (float)acc.sat; v[31] = (float)acc.patches;  // counts < 2^24: exact in fp32
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const float keep = up ? v[j + o] : v[j];
      const float send = up ? v[j] : v[j + o];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  wp[lane < 28 ? lane : lane + 7] = (double)v[0];  // 0..27 -> H, 28 -> E (35), 29..31 -> terms, saturated, patches (36..38)
  // the 7 fp64 gradient entries: three transposed steps (4+2+1 shuffles) then two plain butterfly steps
  double d[8];
#pragma unroll
  for (int k = 0; k < 7; ++k) d[k] = acc.b[k];
  d[7] = 0.0;
#pragma unroll
  for (int o = 16, n = 4; o >= 4; o >>= 1, n >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const double keep = up ? d[j + n] : d[j];
      const double send = up ? d[j] : d[j + n];
      d[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  d[0] += __shfl_xor_sync(0xffffffffu, d[0], 2);
  d[0] += __shfl_xor_sync(0xffffffffu, d[0], 1);
  if ((lane & 3) == 0 && (lane >> 2) < 7) wp[28 + (lane >> 2)] = d[0];
  __syncthreads();
  if (threadIdx.x < NRED - 1) {
    double sum = 0;
    for (int w = 0; w < nwarps; ++w) sum += s.warp_part[w * NRED + threadIdx.x];
    if (csize == 1) s.tot[threadIdx.x] = sum;
    else s.cta_part[slot * NRED + threadIdx.x] = sum;
  }
  if (csize == 1) {
    __syncthreads();
    return;
  }
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  if (threadIdx.x < NRED - 1) {
    double sum = 0;
    for (int r = 0; r < csize; ++r) {
      const double* peer = cluster.map_shared_rank(s.cta_part, r);
      sum += peer[slot * NRED + threadIdx.x];
    }
    s.tot[threadIdx.x] = sum;
  }
  __syncthreads();
}

// Merge the histograms of the cluster's CTAs (DSMEM) and locate the bucket that holds rank k (k = total / 2 when `first`: hso::getMedian,
// include/hso/vikit/math_utils.h:119-126, else ctrl->sel_k). Block-parallel: each thread sums `per` consecutive bins, warp scan, warp totals
// through shared memory. Results (identical in every CTA of the cluster): ctrl->sel_bin, ->sel_k (rank inside the bucket), ->sel_cnt (bucket
// population), ->sel_n (total, when first).
HSO_DEV void find_bucket(const Smem& s, int nbins, int csize, bool first) {
  if (csize == 1) {
    __syncthreads();
  } else {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    for (int j = threadIdx.x; j < nbins; j += blockDim.x) {
      uint32_t sum = 0;
      for (int r = 0; r < csize; ++r) sum += cluster.map_shared_rank(s.hist, r)[j];
      s.ghist[j] = sum;
    }
    cluster.sync();  // peers are done reading this CTA's histogram before it is zeroed again
  }
  const uint32_t* gh = csize == 1 ? s.hist : s.ghist;  // csize == 1 is also used by a cluster's CTAs on identical private copies
  const int T = blockDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = T >> 5;
  const int per = (nbins + T - 1) / T;
  const int b0 = threadIdx.x * per;
  uint32_t local = 0;
  for (int j = 0; j < per; ++j) local += (b0 + j < nbins) ? gh[b0 + j] : 0u;
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s.ctrl->wsum[warp] = incl;
  __syncthreads();
  uint32_t before = 0, total = 0;
  for (int q = 0; q < nw; ++q) {
    const uint32_t v = s.ctrl->wsum[q];
    if (q < warp) before += v;
    total += v;
  }
  const uint32_t k = first ? total / 2 : s.ctrl->sel_k;
  const uint32_t excl = before + incl - local;
  __syncthreads();  // everyone has read sel_k / wsum before they are overwritten
  if (first && threadIdx.x == 0) s.ctrl->sel_n = total;
  if (total > 0 && k >= excl && k < excl + local) {
    uint32_t cum = excl;
    int d = b0;
    uint32_t cnt = 0;
    for (int j = 0; j < per; ++j) {
      const uint32_t c = gh[b0 + j];
      if (k < cum + c) { d = b0 + j; cnt = c; break; }
      cum += c;
    }
    s.ctrl->sel_k = k - cum;
    s.ctrl->sel_bin = (uint32_t)d;
    s.ctrl->sel_cnt = cnt;
  }
  if (total == 0 && threadIdx.x == 0) { s.ctrl->sel_k = 0; s.ctrl->sel_bin = 0; s.ctrl->sel_cnt = 0; }
  __syncthreads();
}

// Visit every value f(absres) the CTA owns (MAD: fabsf(v - center)); invalid slots (absres < 0) are skipped. The N values of a patch are
// loaded back to back (one round trip per patch) into N registers — a larger chunk spills at the kernel's 128-register budget. `absres` is a
// generic pointer (shared memory or the L2-resident global scratch): one load instruction serves both layouts.
template <int N, class Fn>
HSO_DEV void for_each_abs(const TrackJobDev& job, const float* absres, int astride, bool a_smem, int t0, int nt, bool mad, float center, Fn fn) {
  int kk = 0;
  for (int i = t0; i < job.F; i += nt, ++kk) {
    const float* p = absres + (a_smem ? kk * (int)blockDim.x + (int)threadIdx.x : i);
    float vals[N];
#pragma unroll
    for (int n = 0; n < N; ++n) vals[n] = p[n * astride];
    if (!(vals[0] >= 0.f)) continue;  // a patch is in view with all of its N pattern pixels or with none (the threshold pass writes -1 to all)
#pragma unroll
    for (int n = 0; n < N; ++n) fn(mad ? fabsf(vals[n] - center) : vals[n]);
  }
}

// Exact k-th smallest (k = n/2) of the non-negative floats { f(absres) : absres >= 0 } owned by the cluster; result in ctrl->sel_prefix
// (the float's bit pattern), population in ctrl->sel_n. Radix select over the IEEE bit pattern (monotone for non-negative floats), digits of
// `bits` bits from the top; prefilled: the histogram of the first digit was already accumulated by the caller (fused with the residuals).
template <int N>
HSO_DEV void radix_select(const TrackJobDev& job, const Smem& s, const float* absres, int astride, bool a_smem, int t0, int nt, bool mad, float center,
                          int csize, int bits, bool prefilled) {
  const int nbins = 1 << bits;
  uint32_t prefix = 0, mask = 0;
  int hi = 32;
  bool first = true;
  while (hi > 0) {
    const int width = hi >= bits ? bits : hi;
    const int shift = hi - width;
    const uint32_t dmask = (1u << width) - 1u;
    uint32_t* hist = s.hist;
    if (!(first && prefilled)) {
      for (int j = threadIdx.x; j < nbins; j += blockDim.x) hist[j] = 0;
      __syncthreads();
      for_each_abs<N>(job, absres, astride, a_smem, t0, nt, mad, center, [&](float v) {
        const uint32_t key = __float_as_uint(v);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & dmask], 1u);
      });
    }
    find_bucket(s, nbins, csize, first);
    prefix |= s.ctrl->sel_bin << shift;
    mask |= dmask << shift;
    hi = shift;
    first = false;
  }
  __syncthreads();
  if (threadIdx.x == 0) s.ctrl->sel_prefix = s.ctrl->sel_n ? prefix : 0u;
  __syncthreads();
}

// Monotone bucket of the first selection pass: 1/64 grey level per bin (values >= 32 share the top bin). The residual magnitudes of a level cluster around a few grey levels,
// so float-exponent digits would put >10 % of all values into one bin (atomics serialise, and the bin's members must be re-scanned twice);
// linear bins of 1/64 grey level spread them (<1 % per bin) and the members of the chosen bin fit a small list.
constexpr float LIN_SCALE = 64.f;
constexpr int LIST_BINS = 512;              // digit width of the list passes
constexpr int LIST_CAP = 2048 - LIST_BINS;  // the list shares the 2048-word histogram buffer with a LIST_BINS-bin histogram
HSO_DEV uint32_t lin_bin(float v, int nbins) {
  const int b = (int)(v * LIN_SCALE);
  return (uint32_t)(b < nbins - 1 ? b : nbins - 1);
}

// Compaction pass of select_kth: the members of linear bin `bin` go to `list`. Branch-free membership mask over the patch's N values, ONE
// (rarely taken: a bin holds < 1 % of the values) branch per patch instead of one per value.
template <int N>
HSO_DEV void compact_bin(const TrackJobDev& job, const float* absres, int astride, bool a_smem, int t0, int nt, bool mad, float center, uint32_t bin,
                         int nbins, uint32_t* list, uint32_t* list_n) {
  int kk = 0;
  for (int i = t0; i < job.F; i += nt, ++kk) {
    const float* p = absres + (a_smem ? kk * (int)blockDim.x + (int)threadIdx.x : i);
    float vals[N];
#pragma unroll
    for (int n = 0; n < N; ++n) vals[n] = p[n * astride];
    if (!(vals[0] >= 0.f)) continue;
    uint32_t hit = 0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      if (mad) vals[n] = fabsf(vals[n] - center);
      hit |= (lin_bin(vals[n], nbins) == bin ? 1u : 0u) << n;
    }
    if (hit) {
#pragma unroll
      for (int n = 0; n < N; ++n)
        if ((hit >> n) & 1u) list[atomicAdd(list_n, 1u)] = __float_as_uint(vals[n]);
    }
  }
}

// Same result as radix_select, fewer full passes: (1) histogram over linear bins (already accumulated when `prefilled`), (2) ONE pass that
// compacts the members of the chosen bin into a list in shared memory — the CTAs of a cluster then copy each other's lists through DSMEM so
// that every CTA holds the whole bin, (3) radix select over that list only, CTA-local (no cluster barriers), on the key's offset from the
// bin's lower bound: a bin of 1/64 grey level spans <= 2^17 bit patterns for v >= 1, i.e. two 9-bit passes. Falls back to the plain radix
// select when the bin is the unbounded top one or too populated for the list (e.g. identical images: every |r| is 0).
template <int N>
HSO_DEV void select_kth(const TrackJobDev& job, const Smem& s, const float* absres, int astride, bool a_smem, int t0, int nt, bool mad, float center,
                        int csize, bool prefilled) {
  constexpr int NB = 2048;
  if (!prefilled) {
    for (int j = threadIdx.x; j < NB; j += blockDim.x) s.hist[j] = 0;
    __syncthreads();
    for_each_abs<N>(job, absres, astride, a_smem, t0, nt, mad, center, [&](float v) { atomicAdd(&s.hist[lin_bin(v, NB)], 1u); });
  }
  find_bucket(s, NB, csize, true);
  const uint32_t total = s.ctrl->sel_n, bin = s.ctrl->sel_bin, cnt = s.ctrl->sel_cnt;
  if (total == 0) {
    if (threadIdx.x == 0) s.ctrl->sel_prefix = 0;
    __syncthreads();
    return;
  }
  if (bin == NB - 1 || cnt > (uint32_t)LIST_CAP) {
    __syncthreads();
    if (threadIdx.x == 0 && cg::this_cluster().block_rank() == 0) job.state->cycles[7] += 1ull << 32;  // diagnostics: fallbacks (upper word)
    radix_select<N>(job, s, absres, astride, a_smem, t0, nt, mad, center, csize, 11, false);
    return;
  }
  // (2) compaction of the bin's members (this CTA's share) behind the LIST_BINS-bin histogram
  uint32_t* list = s.hist + LIST_BINS;
  if (threadIdx.x == 0) s.ctrl->list_n = 0;
  __syncthreads();
  compact_bin<N>(job, absres, astride, a_smem, t0, nt, mad, center, bin, NB, list, &s.ctrl->list_n);
  int ln;
  if (csize == 1) {
    __syncthreads();
    ln = (int)s.ctrl->list_n;
  } else {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();  // every CTA's share is complete
    const int own = (int)s.ctrl->list_n;
    const unsigned me = cluster.block_rank();
    int off = own;
    for (int r = 0; r < csize; ++r) {
      if ((unsigned)r == me) continue;
      const int rn = (int)cluster.map_shared_rank(&s.ctrl->list_n, r)[0];
      const uint32_t* rl = cluster.map_shared_rank(list, r);
      for (int j = threadIdx.x; j < rn; j += blockDim.x) list[off + j] = rl[j];  // appended behind the own share, which peers are reading
      off += rn;
    }
    ln = off;
    cluster.sync();  // peers have read this CTA's share and count: the buffer may be reused after the select
  }
  // (3) rank sel_k inside the bin, on the key's offset from the bin's lower bound
  const uint32_t key_lo = __float_as_uint((float)bin / LIN_SCALE);
  const uint32_t range = __float_as_uint((float)(bin + 1) / LIN_SCALE) - key_lo;  // keys of the bin lie in [key_lo, key_lo + range)
  int hi = 32 - __clz((int)(range - 1));
  if (range <= 1) hi = 0;
  uint32_t prefix = 0, mask = 0;
  while (hi > 0) {
    const int width = hi >= 9 ? 9 : hi;
    const int shift = hi - width;
    const uint32_t dmask = (1u << width) - 1u;
    for (int j = threadIdx.x; j < LIST_BINS; j += blockDim.x) s.hist[j] = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < ln; j += blockDim.x) {
      const uint32_t rel = list[j] - key_lo;
      if ((rel & mask) == prefix) atomicAdd(&s.hist[(rel >> shift) & dmask], 1u);
    }
    find_bucket(s, LIST_BINS, 1, false);
    prefix |= s.ctrl->sel_bin << shift;
    mask |= dmask << shift;
    hi = shift;
  }
  __syncthreads();
  if (threadIdx.x == 0) s.ctrl->sel_prefix = key_lo + prefix;
  __syncthreads();
}

// Slow path of the damped solve: diagonal-pivoted LDL^T with Eigen::LDLT's semantics on a stack copy (dynamic indexing).
__device__ __noinline__ void solve_pivoted(const TrackCtrl* c, float lambda, double* step) {
  double Hl[49], bl[7];
  int idx = 0;
  for (int r = 0; r < 7; ++r)
    for (int q = r; q < 7; ++q) { Hl[r * 7 + q] = Hl[q * 7 + r] = c->H[idx]; ++idx; }
  for (int k = 0; k < 7; ++k) Hl[k * 7 + k] *= (double)(1.f + lambda);
  for (int k = 0; k < 7; ++k) bl[k] = c->b[k];
  ldlt_solve<7>(Hl, bl, step);
}

// Thread-0 control step after a residual evaluation: accept/reject, damping update, convergence test, damped 7x7 solve and
// SE3 update for the next trial (src/CoarseTracker.cpp:108-194). Every CTA of a cluster runs it redundantly on identical totals.
// (The job's pointers come by value: a reference to the kernel's local TrackJobDev copy would pin that copy to the local-memory stack, and every
// control step would start with dependent loads that miss the small L1 left beside 180-230 KB of shared memory.)
__device__ __noinline__ void lm_control(TrackCtrl* c, const double* tot, hso_trace* job_trace, TrackState* job_state, int iter, int n_iter, int level,
                                        int trace_cap, bool ic, bool rank0, float a_eval, float huber, float cutoff, int N) {
  const int terms = (int)tot[36];
  const double E = (double)((float)tot[35] / (float)terms);  // return E/m_total_terms (float / int), :413
  const bool accepted = iter < 0 ? true : (E < c->E_old);
  if (job_trace != nullptr && rank0) {
    const int tl = job_state->trace_len;
    if (tl < trace_cap) {
      hso_trace* e = job_trace + tl;
      e->level = level; e->iter = iter;
      for (int k = 0; k < 12; ++k) e->T_eval[k] = c->Rt[k];
      e->a_eval = a_eval; e->lambda = iter < 0 ? 0.f : c->lambda;
      int idx = 0;
      for (int r = 0; r < 7; ++r)
        for (int q = r; q < 7; ++q) { e->H[r * 7 + q] = e->H[q * 7 + r] = tot[idx]; ++idx; }
      for (int k = 0; k < 7; ++k) { e->b[k] = tot[28 + k]; e->step[k] = iter < 0 ? 0.0 : c->step[k]; }
      e->energy = E; e->total_terms = terms; e->saturated_terms = (int)tot[37];
      e->accepted = accepted ? 1 : 0; e->huber = huber; e->outlier = cutoff;
      job_state->trace_len = tl + 1;
    }
  }
  if (iter < 0) c->lambda = 0.1f;
  // The system the next solve uses: the one just evaluated when it is accepted, else the stored one. All 35 values are loaded into
  // registers first (independent loads), then stored: a load/store-interleaved copy between two shared-memory pointers the compiler cannot
  // prove distinct is a chain of 35 dependent round trips (measured: 1.1 k of the control step's 4.5 k cycles).
  double Hd[28], bd[7];
  {
    const double* srcH = accepted ? tot : c->H;
    const double* srcb = accepted ? tot + 28 : c->b;
#pragma unroll
    for (int k = 0; k < 28; ++k) Hd[k] = srcH[k];
#pragma unroll
    for (int k = 0; k < 7; ++k) bd[k] = srcb[k];
  }
  if (accepted) {
#pragma unroll
    for (int k = 0; k < 28; ++k) c->H[k] = Hd[k];
#pragma unroll
    for (int k = 0; k < 7; ++k) c->b[k] = bd[k];
    c->E_old = E;
    if (iter >= 0) {
      c->a_acc = c->a_try;
      c->T_acc = c->T_try;
      c->lambda *= 0.5f;
    }
  } else {
    c->lambda *= 4.f;
    if (c->lambda < 0.001f) c->lambda = 0.001f;
  }
  bool done = false;
  if (iter >= 0) {
    double nrm = 0;
    for (int k = 0; k < 7; ++k) nrm += c->step[k] * c->step[k];
    if (!(sqrt(nrm) > 1e-4)) done = true;  // :188
  }
  if (iter + 1 >= n_iter) done = true;
  if (rank0) { job_state->last_total_terms = terms; job_state->last_N = N; }
  if (!done) {
    // damped solve, extrapolation, NaN guard (:112-124)
    double step[7];
    const float lambda = c->lambda;
    bool solved;
    {
      // register-resident unpivoted factorisation when the damped system is safely positive definite (the normal case). Every index below is
      // a compile-time constant and no address escapes, so Hl / bl live in registers (a stack copy would be read back through an L1 that is
      // almost entirely carved out as shared memory).
      double Hl[49], bl[7];
      int idx = 0;
#pragma unroll
      for (int r = 0; r < 7; ++r)
#pragma unroll
        for (int q = r; q < 7; ++q) { Hl[r * 7 + q] = Hl[q * 7 + r] = Hd[idx]; ++idx; }
#pragma unroll
      for (int k = 0; k < 7; ++k) Hl[k * 7 + k] *= (double)(1.f + lambda);
#pragma unroll
      for (int k = 0; k < 7; ++k) bl[k] = bd[k];
      solved = ldlt_solve_spd_fast<7>(Hl, bl, step);
    }
    if (!solved) {  // the pivoted robust-Cholesky path (Eigen::LDLT semantics)
      solve_pivoted(c, lambda, step);
      if (rank0) job_state->cycles[7] += 1;  // diagnostics: number of pivoted (slow path) solves
    }
    float extrap = 1.f;
    if (lambda < 0.001f) extrap = (float)sqrt(sqrt(0.001 / (double)lambda));
    double sum = 0;
    for (int k = 0; k < 7; ++k) { step[k] *= (double)extrap; sum += step[k]; }
    if (!isfinite(sum) || isnan(step[0])) for (int k = 0; k < 7; ++k) step[k] = 0.0;
    for (int k = 0; k < 7; ++k) c->step[k] = step[k];
    c->a_try = (float)((double)c->a_acc + step[0]);
    double neg[6];
    for (int k = 0; k < 6; ++k) neg[k] = -step[1 + k];
    const Se3d dT = se3_exp(neg);
    c->T_try = ic ? se3_mul(c->T_acc, dT) : se3_mul(dT, c->T_acc);
    se3_to_rt(c->T_try, c->Rt);
  }
  c->done = done ? 1 : 0;
}

template <int PIDX, bool IC, int MODE>
__global__ void __launch_bounds__(512, 1) k_track_level(const TrackLevelParams prm, const TrackJobDev* __restrict__ jobs) {
  constexpr int N = (PIDX == 2) ? 9 : (PIDX == 3 || PIDX == 4) ? 13 : (PIDX == 5) ? 21 : 25;
  constexpr int PAD = (PIDX == 5) ? 3 : (PIDX == 7) ? 4 : (PIDX <= 2) ? 1 : 2;
  constexpr bool FAST = MODE != 0, DUAL = MODE == 2, STREAM = MODE >= 3;
  constexpr int NR = IC ? 3 * N : N;
  constexpr int RD = MODE == 4 ? 1 : 2;
  static_assert(!DUAL || IC, "the dual-image mode exists for the inverse-compositional path only");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const int problem = blockIdx.x / csize;
  const TrackJobDev job = jobs[problem];
  const int nwarps = blockDim.x >> 5;
  const long long clk_start = clock64();

  Smem s;
  {
    size_t oc, oa, ov, ow, op, ot, oh, og, ox, om;
    const size_t abs_bytes = (FAST && prm.absres_smem) ? (size_t)N * prm.pc * sizeof(float) : 0;
    const size_t cache_bytes = MODE == 1 ? (size_t)N * (IC ? 3 : 1) * prm.pc * sizeof(float) : STREAM ? ring_bytes(NR, nwarps, RD) : 0;
    const uint32_t img_total = DUAL ? 2 * (uint32_t)align_up(prm.img_bytes, 128) : (FAST ? prm.img_bytes : 0);
    smem_layout(img_total, cache_bytes, abs_bytes, FAST ? (size_t)prm.pc : 0, nwarps, prm.hist_bits, csize, &oc, &oa, &ov, &ow, &op, &ot, &oh, &og, &ox, &om);
    s.absres = reinterpret_cast<float*>(smem_raw + oa);
    s.img = smem_raw;
    s.cache = reinterpret_cast<float*>(smem_raw + oc);
    s.vis = smem_raw + ov;
    s.warp_part = reinterpret_cast<double*>(smem_raw + ow);
    s.cta_part = reinterpret_cast<double*>(smem_raw + op);
    s.tot = reinterpret_cast<double*>(smem_raw + ot);
    s.hist = reinterpret_cast<uint32_t*>(smem_raw + oh);
    s.ghist = reinterpret_cast<uint32_t*>(smem_raw + og);
    s.ctrl = reinterpret_cast<TrackCtrl*>(smem_raw + ox);
    s.mbar = reinterpret_cast<uint64_t*>(smem_raw + om);
  }
  TrackCtrl* c = s.ctrl;
  const int t0 = crank * blockDim.x + threadIdx.x;
  const int nt = csize * blockDim.x;
  const int Fp = job.Fpad;

  PatchStore ps;
  ps.ref = s.img + align_up(prm.img_bytes, 128);
  ps.ring_bar = reinterpret_cast<uint64_t*>(s.cache) + 2 * (threadIdx.x >> 5);
  ps.ring = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s.cache) + ring_bar_bytes(nwarps)) + (threadIdx.x >> 5) * (RD * NR * 32);
  uint32_t ring_g = 0;  // groups this warp has streamed so far (MODE 3)
  if (STREAM) {
    ps.cache = job.ref_cache; ps.gx = nullptr; ps.gy = nullptr; ps.vis = s.vis; ps.stride = Fp;
    if ((threadIdx.x & 31) == 0) { mbar_init(ps.ring_bar, 1); mbar_init(ps.ring_bar + 1, 1); mbar_fence_init(); }
  } else if (FAST) {
    ps.cache = s.cache; ps.gx = s.cache + (size_t)N * prm.pc; ps.gy = s.cache + (size_t)2 * N * prm.pc; ps.vis = s.vis; ps.stride = prm.pc;
  } else {
    ps.cache = job.ref_cache; ps.gx = job.ref_gx; ps.gy = job.ref_gy; ps.vis = job.vis; ps.stride = Fp;
  }

  // ---- phase 0: stage the current level image (TMA bulk copy), read the accepted state ---------------------------------
  const uint8_t* cur_g = job.cur_pyr + prm.level_off;
  const uint8_t* ref_g = job.ref_pyr + prm.level_off;
  if (threadIdx.x == 0) {
    if (FAST) {
      // the staging buffer first holds the REFERENCE level (phase 1 gathers the reference patches from it), then the CURRENT level
      mbar_init(s.mbar, 1);
      mbar_fence_init();
      fence_proxy_async();
      mbar_expect_tx(s.mbar, DUAL ? 2 * prm.img_bytes : prm.img_bytes);
      uint32_t done = 0;
      while (done < prm.img_bytes) {
        uint32_t chunk = prm.img_bytes - done;
        if (chunk > 32768u) chunk = 32768u;
        if (DUAL) {  // both levels stay resident: current at s.img, reference behind it
          tma_bulk_g2s(s.img + done, cur_g + done, chunk, s.mbar);
          tma_bulk_g2s(s.img + align_up(prm.img_bytes, 128) + done, ref_g + done, chunk, s.mbar);
        } else {
          tma_bulk_g2s(s.img + done, ref_g + done, chunk, s.mbar);
        }
        done += chunk;
      }
    }
    c->T_acc = job.state->T;
    c->a_acc = job.state->a;
    c->done = 0;
    c->iter = -1;
  }

  LevelCtx L;
  L.cur = FAST ? s.img : cur_g;
  L.w = prm.w; L.h = prm.h; L.border = PAD + 1;
  L.scale = 1.0f / (float)(1 << prm.level);
  L.fxl = (float)(prm.cam.fx * (double)L.scale);
  L.fyl = (float)(prm.cam.fy * (double)L.scale);
  L.top = prm.level == prm.max_level;

  // ---- phase 1: precomputeReferencePatches (src/CoarseTracker.cpp:416-497) ------------------------------------------------
  __syncthreads();
  if (FAST) mbar_wait(s.mbar, 0);
  const uint8_t* ref_src = DUAL ? ps.ref : (FAST ? s.img : ref_g);
  {
    int k = 0;
    for (int i = t0; i < job.F; i += nt, ++k) {
      const int sl = slot_of<FAST>(i, k);

      const RefPatch rp = ref_patch(job.px[i], job.px[Fp + i], L.scale, L.border, L.w, L.h);
      ps.vis[sl] = rp.in ? 1 : 0;
      if (!rp.in || DUAL) continue;  // dual-image mode recomputes the reference samples in every evaluation
      if (!IC) {
        window_colors<PIDX, FAST>(ref_src, rp.base, L.w, rp.wtl, rp.wtr, rp.wbl, rp.wbr, [&](int n, float cc) {
          ps.cache[STREAM ? ring_index<NR>(i, n) : n * ps.stride + sl] = cc;
        });
        continue;
      }
      // inverse-compositional: intensity and central-difference gradients of the reference image (:482-492) from one streamed window
      window_samples<PIDX, FAST>(ref_src, rp.base, L.w, rp.wtl, rp.wtr, rp.wbl, rp.wbr, [&](int n, float cc, float gx2, float gy2) {
        const float gx = 0.5f * gx2, gy = 0.5f * gy2;
        if (STREAM) {
          ps.cache[ring_index<NR>(i, n)] = cc;
          ps.cache[ring_index<NR>(i, N + n)] = gx;
          ps.cache[ring_index<NR>(i, 2 * N + n)] = gy;
        } else {
          ps.cache[n * ps.stride + sl] = cc;
          ps.gx[n * ps.stride + sl] = gx;
          ps.gy[n * ps.stride + sl] = gy;
        }
      });
    }
  }
  // MODE 3: the cache just written through the generic proxy is read back by TMA bulk copies (async proxy)
  if (STREAM) asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncthreads();
  const long long clk_pre = clock64();
  if (FAST && !DUAL) {
    // every generic-proxy read of the reference image is done: hand the buffer back to the async proxy for the current level
    if (threadIdx.x == 0) {
      fence_proxy_async();
      mbar_expect_tx(s.mbar, prm.img_bytes);
      uint32_t done = 0;
      while (done < prm.img_bytes) {
        uint32_t chunk = prm.img_bytes - done;
        if (chunk > 32768u) chunk = 32768u;
        tma_bulk_g2s(s.img + done, cur_g + done, chunk, s.mbar);
        done += chunk;
      }
    }
    mbar_wait(s.mbar, 1);
  }
  if (threadIdx.x == 0) se3_to_rt(c->T_acc, c->Rt);
  __syncthreads();

  // ---- phase 2: selectRobustFunctionLevel (src/CoarseTracker.cpp:530-644) ------------------------------------------------
  long long clk_res = 0, clk_med = 0;
  {
    const float a = c->a_acc;
    const int hbits = prm.hist_bits;
    const bool a_smem = FAST && prm.absres_smem;
    const bool lin = hbits == 11;  // the 2048-word buffer holds the linear histogram, then a 256-bin histogram + the candidate list
    float* absres = a_smem ? s.absres : job.absres;
    const int astride = a_smem ? prm.pc : Fp;
    for (int j = threadIdx.x; j < (1 << hbits); j += blockDim.x) s.hist[j] = 0;
    __syncthreads();
    int k = 0;
    const int lane = threadIdx.x & 31;
    if (STREAM && t0 - lane < job.F) ring_issue<NR, RD>(ps, t0 - lane, ring_g);
    for (int i = t0; (STREAM ? i - lane : i) < job.F; i += nt, ++k) {  // (STREAM: warp-uniform trip count, the ring is filled by the whole warp)
      const float* cbuf = nullptr;
      if (STREAM) {
        if (RD == 2 && i + nt - lane < job.F) ring_issue<NR, RD>(ps, i + nt - lane, ring_g + 1);
        cbuf = ring_wait<NR, RD>(ps, ring_g);
      }
      const int sl = slot_of<FAST>(i, k);
      bool ok = (!STREAM || i < job.F) && ps.vis[sl] != 0;
      Proj p;
      RefPatch rp;
      if (ok) {
        p = project_patch(c->Rt, prm.cam, job.xyz[i], job.xyz[Fp + i], job.xyz[2 * Fp + i], L.scale, L.border, L.w, L.h);
        ok = p.ok;
        if (DUAL) rp = ref_patch(job.px[i], job.px[Fp + i], L.scale, L.border, L.w, L.h);
      }
      if (ok) {
        // |residual| of every pattern pixel at the level's initial state (:557-606), the current colour from the streamed window
        window_colors<PIDX, FAST>(L.cur, p.base, L.w, p.wtl, p.wtr, p.wbl, p.wbr, [&](int n, float color) {
          const float cref = DUAL ? ref_intensity<true>(ps.ref, rp, rp.base + pat_dy<PIDX>(n) * L.w + pat_dx<PIDX>(n), L.w)
                                  : STREAM ? cbuf[n * 32] : ps.cache[n * ps.stride + sl];
          const float out = fabsf(fmaf(-a, cref, color));
          atomicAdd(&s.hist[lin ? lin_bin(out, 2048) : (__float_as_uint(out) >> (32 - hbits))], 1u);  // first pass of the median select, fused
          absres[n * astride + (a_smem ? sl : i)] = out;
        });
      } else if (!STREAM || i < job.F) {
#pragma unroll
        for (int n = 0; n < N; ++n) absres[n * astride + (a_smem ? sl : i)] = -1.f;
      }
      if (STREAM) {
        __syncwarp();
        if (RD != 2 && i + nt - lane < job.F) ring_issue<NR, RD>(ps, i + nt - lane, ring_g + 1);
        ++ring_g;
      }
    }
    clk_res = clock64();
    if (lin) select_kth<N>(job, s, absres, astride, a_smem, t0, nt, false, 0.f, csize, true);
    else radix_select<N>(job, s, absres, astride, a_smem, t0, nt, false, 0.f, csize, hbits, true);
    clk_med = clock64();
    const uint32_t n_err = c->sel_n;
    float huber = 5.2f, outlier = 100.f;
    if (n_err >= 30) {
      const float median = __uint_as_float(c->sel_prefix);
      __syncthreads();
      if (lin) select_kth<N>(job, s, absres, astride, a_smem, t0, nt, true, median, csize, false);
      else radix_select<N>(job, s, absres, astride, a_smem, t0, nt, true, median, csize, hbits, false);
      const float mad = __uint_as_float(c->sel_prefix);
      const float sd = (float)(1.4826 * (double)mad);
      huber = median + sd;
      outlier = 3.f * huber;
      if (outlier < 10.f) outlier = 10.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) { c->huber = huber; c->outlier = outlier; c->n_err = (int)n_err; }
    __syncthreads();
  }
  const long long clk_lm = clock64();

  // ---- phase 3: the LM loop of the level (src/CoarseTracker.cpp:100-194) -------------------------------------------------
  const float huber = c->huber, cutoff = c->outlier;
  int slot = 0;
  int level_iters = 0, level_evals = 0;
  unsigned long long patch_evals = 0;
  long long clk_ctrl = 0;
  for (int iter = -1; iter < prm.n_iter; ++iter) {
    const float a_eval = iter < 0 ? c->a_acc : c->a_try;
    Acc acc;
    acc_zero(acc);
    // the top level has its own energy and no cut-off (:350-361). m_offset_all = m_max_level - m_level + m_pattern_offset (:80) makes the top
    // level the one and only user of pattern index m_pattern_offset = 2: a compile-time property of the instantiation
    eval_patches<PIDX, IC, MODE, PIDX == 2>(L, job, ps, prm.cam, c->Rt, a_eval, huber, cutoff, t0, nt, acc, ring_g);
    reduce_acc(acc, s, slot, csize, nwarps);
    slot ^= 1;
    ++level_evals;
    if (iter >= 0) ++level_iters;
    patch_evals += (unsigned long long)s.tot[38];
    if (threadIdx.x == 0) {
      const long long t_in = clock64();
      lm_control(c, s.tot, job.trace, job.state, iter, prm.n_iter, prm.level, prm.trace_cap, IC, crank == 0, a_eval, huber, cutoff, N);
      clk_ctrl += clock64() - t_in;
    }
    __syncthreads();
    if (c->done) break;
  }

  // ---- write the accepted state back -------------------------------------------------------------------------------------
  if (threadIdx.x == 0 && crank == 0) {
    TrackState* st = job.state;
    st->T = c->T_acc;
    st->a = c->a_acc;
    st->n_iters += level_iters;
    st->n_evals += level_evals;
    st->iters_per_level[prm.level & 7] = level_iters;
    st->patch_evals[prm.level & 7] = patch_evals;
    const long long clk_end = clock64();
    st->cycles[0] += (unsigned long long)(clk_end - clk_start);  // whole launch (CTA rank 0)
    st->cycles[1] += (unsigned long long)(clk_lm - clk_start);   // staging + reference patches + robust thresholds
    st->cycles[2] += (unsigned long long)clk_ctrl;               // thread-0 control (solve, SE3 update, trace)
    st->cycles[3] += (unsigned long long)(clk_pre - clk_start);  // staging of the reference level + reference patches
    st->cycles[4] += (unsigned long long)(clk_res - clk_pre);    // staging of the current level + residuals of the threshold selection
    st->cycles[5] += (unsigned long long)(clk_med - clk_res);    // median select
    st->cycles[6] += (unsigned long long)(clk_lm - clk_med);     // MAD select
  }
  if (csize > 1) cluster.sync();  // peers may still be reading this CTA's shared memory
}

__global__ void k_track_init(const TrackJobDev* __restrict__ jobs, const double* __restrict__ T0, const float* __restrict__ a0, int B) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B) return;
  TrackState* st = jobs[p].state;
  st->T = se3_from_rt(T0 + 12 * p);
  // float a = cur.integralImage_/ref.integralImage_ (src/CoarseTracker.cpp:60) when the caller left it to the device
  st->a = a0[p] < 0.f ? jobs[p].cur_stats[0] / jobs[p].ref_stats[0] : a0[p];
  st->n_iters = 0; st->n_evals = 0;
  for (int k = 0; k < 8; ++k) { st->iters_per_level[k] = 0; st->patch_evals[k] = 0; }
  st->last_total_terms = 0; st->last_N = 1; st->trace_len = 0;
  for (int k = 0; k < 8; ++k) st->cycles[k] = 0;
}

__global__ void k_track_finish(const TrackJobDev* __restrict__ jobs, hso_track_result* __restrict__ out, int B) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B) return;
  const TrackState* st = jobs[p].state;
  hso_track_result* o = out + p;
  se3_to_rt(st->T, o->T_cur_ref);
  o->exposure_rat = st->a;
  o->n_iters = st->n_iters;
  o->n_evals = st->n_evals;
  for (int k = 0; k < 8; ++k) { o->iters_per_level[k] = st->iters_per_level[k]; o->visible_patch_evals[k] = st->patch_evals[k]; }
  // return float(m_total_terms) / PATCH_AREA  (src/CoarseTracker.cpp:207)
  o->n_tracked = (uint64_t)((float)st->last_total_terms / (float)st->last_N);
  o->trace_len = st->trace_len;
  for (int k = 0; k < 8; ++k) o->cycles[k] = st->cycles[k];
}

// Direct-input mode: what track_stage_one does on the host (capi.cu), on the device. One CTA per job; features with a valid depth (dist >= 0)
// are kept in order — a block-wide prefix scan per tile of 256 features — and xyz = f * dist is the same single IEEE fp64 multiplication as
// Vector3d xyz_ref((*it_ft)->f*dist) (src/CoarseTracker.cpp:292): bit-identical to the host path.
__global__ void __launch_bounds__(256) k_track_compact(TrackJobDev* __restrict__ jobs, int B) {
  if ((int)blockIdx.x >= B) return;
  const TrackJobDev job = jobs[blockIdx.x];
  const int n = job.n_raw, Fp = job.Fpad;
  double* px = const_cast<double*>(job.px);
  double* xyz = const_cast<double*>(job.xyz);
  if (job.raw_xyz != nullptr) {
    // compact layout: xyz formed by the caller, px as float32 ((double)(float)px * 2^-l rounds to the same float as px * 2^-l); all valid
    for (int i = (int)threadIdx.x; i < n; i += 256) {
      px[i] = (double)job.raw_px32[2 * i]; px[Fp + i] = (double)job.raw_px32[2 * i + 1];
      xyz[i] = job.raw_xyz[3 * i]; xyz[Fp + i] = job.raw_xyz[3 * i + 1]; xyz[2 * Fp + i] = job.raw_xyz[3 * i + 2];
    }
    for (int k = n + (int)threadIdx.x; k < Fp; k += 256) { px[k] = 0; px[Fp + k] = 0; xyz[k] = 0; xyz[Fp + k] = 0; xyz[2 * Fp + k] = 1; }
    if (threadIdx.x == 0) jobs[blockIdx.x].F = n;
    return;
  }
  __shared__ int s_warp[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = 0;
  for (int tile = 0; tile < n; tile += 256) {
    const int i = tile + (int)threadIdx.x;
    double d = -1.0;
    if (i < n) d = job.raw_dist[i];
    const bool valid = d >= 0;  // NaN and negative distances are dropped, like if(!(d >= 0)) continue on the host
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int within = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const int c = s_warp[w]; if (w < warp) before += c; total += c; }
    if (valid) {
      const int k = base + before + within;
      px[k] = job.raw_px[2 * i]; px[Fp + k] = job.raw_px[2 * i + 1];
      xyz[k] = job.raw_f[3 * i] * d; xyz[Fp + k] = job.raw_f[3 * i + 1] * d; xyz[2 * Fp + k] = job.raw_f[3 * i + 2] * d;
    }
    base += total;
    __syncthreads();
  }
  for (int k = base + (int)threadIdx.x; k < Fp; k += 256) { px[k] = 0; px[Fp + k] = 0; xyz[k] = 0; xyz[Fp + k] = 0; xyz[2 * Fp + k] = 1; }
  if (threadIdx.x == 0) jobs[blockIdx.x].F = base;
}

cudaError_t launch_track_compact(TrackJobDev* jobs_dev, int B, cudaStream_t stream, uint64_t* launches) {
  k_track_compact<<<B, 256, 0, stream>>>(jobs_dev, B);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_track_init(const TrackJobDev* jobs_dev, const double* T0, const float* a0, int B, cudaStream_t stream, uint64_t* launches) {
  k_track_init<<<(B + 127) / 128, 128, 0, stream>>>(jobs_dev, T0, a0, B);
  ++*launches;
  return cudaGetLastError();
}
cudaError_t launch_track_finish(const TrackJobDev* jobs_dev, hso_track_result* out_dev, int B, cudaStream_t stream, uint64_t* launches) {
  k_track_finish<<<(B + 127) / 128, 128, 0, stream>>>(jobs_dev, out_dev, B);
  ++*launches;
  return cudaGetLastError();
}

template <int PIDX, bool IC, int MODE>
static cudaError_t launch_one(const TrackLevelParams& p, const TrackJobDev* jobs_dev, int B, int cluster, int threads, size_t smem,
                              cudaStream_t stream) {
  auto kern = k_track_level<PIDX, IC, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // all of the unified L1 / shared memory as shared memory: two CTAs of a small-footprint shape (inverse-compositional dual-image mode at the
  // coarse levels) can then share an SM
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  if (getenv("HSO_TRACK_DEBUG")) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem);
    fprintf(stderr, "k_track_level<%d,%d,%d> level shape: cluster %d x %d threads, %zu B smem -> %d CTA(s)/SM\n", PIDX, (int)IC, MODE, cluster, threads, smem, nb);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * cluster), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p, jobs_dev);
}

template <int PIDX>
static cudaError_t launch_pidx(const TrackLevelParams& p, const TrackJobDev* jobs_dev, int B, int cluster, int threads, size_t smem,
                               cudaStream_t stream) {
  if (p.ic) {
    if (p.fast == 2) return launch_one<PIDX, true, 2>(p, jobs_dev, B, cluster, threads, smem, stream);
    if (p.fast == 3) return launch_one<PIDX, true, 3>(p, jobs_dev, B, cluster, threads, smem, stream);
    if (p.fast == 4) return launch_one<PIDX, true, 4>(p, jobs_dev, B, cluster, threads, smem, stream);
    return p.fast ? launch_one<PIDX, true, 1>(p, jobs_dev, B, cluster, threads, smem, stream)
                  : launch_one<PIDX, true, 0>(p, jobs_dev, B, cluster, threads, smem, stream);
  }
  if (p.fast == 3) return launch_one<PIDX, false, 3>(p, jobs_dev, B, cluster, threads, smem, stream);
  return p.fast ? launch_one<PIDX, false, 1>(p, jobs_dev, B, cluster, threads, smem, stream)
                : launch_one<PIDX, false, 0>(p, jobs_dev, B, cluster, threads, smem, stream);
}

cudaError_t launch_track_level(const TrackLevelParams& p, const TrackJobDev* jobs_dev, int B, int cluster, int threads, cudaStream_t stream,
                               uint64_t* launches) {
  const int pidx = p.max_level - p.level + 2;  // m_offset_all (src/CoarseTracker.cpp:80, CoarseTracker.h:122)
  if (pidx < 2 || pidx > 7) return cudaErrorInvalidValue;
  const size_t smem = track_level_smem_bytes(p, threads);
  ++*launches;
  switch (pidx) {
    case 2: return launch_pidx<2>(p, jobs_dev, B, cluster, threads, smem, stream);
    case 3: return launch_pidx<3>(p, jobs_dev, B, cluster, threads, smem, stream);
    case 4: return launch_pidx<4>(p, jobs_dev, B, cluster, threads, smem, stream);
    case 5: return launch_pidx<5>(p, jobs_dev, B, cluster, threads, smem, stream);
    case 6: return launch_pidx<6>(p, jobs_dev, B, cluster, threads, smem, stream);
    default: return launch_pidx<7>(p, jobs_dev, B, cluster, threads, smem, stream);
  }
}

}  // namespace hso
