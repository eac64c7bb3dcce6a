// C-ABI of hso_b200 (include/hso_b200.h): context, device-resident frame table, staging of flattened feature arrays,
// kernel sequencing. Host-side only; all arithmetic of the path lives in the kernels. There is no CPU fallback.
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <atomic>
#include <cmath>
#include <cstddef>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "hso_internal.h"

using namespace hso;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Pinned staging for small records that asynchronous entry points re-write on every call: three regions used in rotation, each guarded by an
// event recorded behind the H2D copy that reads it, so a call never overwrites records a previous call's copy has not consumed yet.
struct PinRing {
  static constexpr int N = 3;
  PinBuf buf;
  size_t slot_cap = 0;
  cudaEvent_t ev[N] = {nullptr, nullptr, nullptr};
  bool pending[N] = {false, false, false};
  int next = 0, cur = 0;
  cudaError_t acquire(size_t bytes, void** out) {
    cudaError_t e;
    if (bytes > slot_cap) {
      for (int i = 0; i < N; ++i)
        if (pending[i]) { if ((e = cudaEventSynchronize(ev[i])) != cudaSuccess) return e; pending[i] = false; }
      const size_t want = (bytes + bytes / 4 + 255) / 256 * 256;
      if ((e = buf.reserve(want * N)) != cudaSuccess) return e;
      slot_cap = want;
    }
    cur = next;
    next = (next + 1) % N;
    if (!ev[cur] && (e = cudaEventCreateWithFlags(&ev[cur], cudaEventDisableTiming)) != cudaSuccess) return e;
    if (pending[cur]) { if ((e = cudaEventSynchronize(ev[cur])) != cudaSuccess) return e; pending[cur] = false; }
    *out = (char*)buf.p + slot_cap * cur;
    return cudaSuccess;
  }
  cudaError_t commit(cudaStream_t stream) {  // call right behind the copy that reads the acquired region
    cudaError_t e = cudaEventRecord(ev[cur], stream);
    if (e == cudaSuccess) pending[cur] = true;
    return e;
  }
  void release() {
    for (int i = 0; i < N; ++i) if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; pending[i] = false; }
    buf.release(); slot_cap = 0;
  }
};

// Host threads that flatten Feature lists to the device layout (track_stage_one). They live as long as the context (spawning eight threads per
// call cost ~0.3 ms and, with one process per GPU, oversubscribed the host: 8 ranks x 8 workers on 32 cores). Size: HSO_FLATTEN_THREADS, else the
// host's cores divided by the ranks sharing it (LOCAL_WORLD_SIZE, as torchrun exports it), capped at 8.
class WorkerPool {
 public:
  typedef void (*Fn)(void* arg, int item);
  ~WorkerPool() { stop(); }
  int size() const { return (int)threads_.size(); }
  void start(int n) {
    if (!threads_.empty() || n <= 0) return;
    for (int t = 0; t < n; ++t) threads_.emplace_back([this]() { loop(); });
  }
  void stop() {
    { std::lock_guard<std::mutex> lk(m_); quit_ = true; }
    cv_.notify_all();
    for (auto& th : threads_) th.join();
    threads_.clear();
    quit_ = false;
  }
  // Runs fn(arg, i) for i in [0, n) on the pool, items handed out in order; returns at once. wait() blocks until all are done.
  void submit(Fn fn, void* arg, int n) {
    { std::lock_guard<std::mutex> lk(m_); fn_ = fn; arg_ = arg; n_ = n; next_.store(0); done_.store(0); ++gen_; }
    cv_.notify_all();
  }
  void wait() {
    // the caller helps: with no threads in the pool it simply runs everything inline
    for (;;) {
      const int i = next_.fetch_add(1, std::memory_order_relaxed);
      if (i >= n_) break;
      fn_(arg_, i);
      done_.fetch_add(1, std::memory_order_release);
    }
    while (done_.load(std::memory_order_acquire) < n_) std::this_thread::yield();
  }

 private:
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&]() { return quit_ || gen_ != seen; });
        if (quit_) return;
        seen = gen_;
      }
      for (;;) {
        const int i = next_.fetch_add(1, std::memory_order_relaxed);
        if (i >= n_) break;
        fn_(arg_, i);
        done_.fetch_add(1, std::memory_order_release);
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_;
  bool quit_ = false;
  uint64_t gen_ = 0;
  Fn fn_ = nullptr;
  void* arg_ = nullptr;
  int n_ = 0;
  std::atomic<int> next_{0}, done_{0};
};

inline int flatten_threads_default() {
  if (const char* e = getenv("HSO_FLATTEN_THREADS")) return std::max(0, std::min(64, atoi(e)));
  int ranks = 1;
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
  const int cores = (int)std::max(1u, std::thread::hardware_concurrency());
  // one core stays with the rank's main thread (it issues the copies and launches)
  return std::max(1, std::min(8, cores / ranks - 1));
}

struct FrameSlot {
  bool used = false;
  uint8_t* pyr = nullptr;
  int16_t* sobel = nullptr;
  double* sums = nullptr;
  float* stats = nullptr;  // device [2]
};

}  // namespace

// Stage timers, named like the reference's HSO_START_TIMER sites (src/frame_handler_base.cpp:57-66) where one exists.
enum { ST_PYRAMID = 0, ST_SPARSE_ALIGN = 1, ST_FEATURE_ALIGN = 2, ST_POSE_OPT = 3, ST_REPROJECT = 4, ST_DEPTH_FILTER = 5, ST_FEATURE_DETECT = 6, kStages = 7 };
static const char* const kStageNames[kStages] = {"pyramid_creation", "sparse_img_align", "feature_align", "pose_optimizer", "reproject",
                                                 "depth_filter_update", "feature_detection"};

struct hso_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  hso_cam cam;
  CamDev camdev;
  hso_cfg cfg;
  PyrGeom geom;
  size_t sobel_elems = 0;
  int n_tiles = 0;
  std::string err;
  uint64_t launches = 0;
  std::vector<FrameSlot> frames;
  size_t free_lo = 0;  // no free slot below this index: allocation returns the LOWEST free slot without rescanning the used prefix
  // pyramid
  DevBuf pyr_arena, sums_arena;  // [max_frames] pyramids (slot stride = pyr_slot_bytes) and per-tile partial sums
  size_t pyr_slot_bytes = 0;
  DevBuf pyr_jobs_dev, pyr_counters, resize_tab_dev, stats_table;  // stats_table: [max_frames][2] floats, one D2H per read
  PinBuf pyr_jobs_host, stats_host;
  PinRing pyr_jobs_ring, t_jobs_ring;  // job records of the asynchronous (unsynchronised) entry points
  WorkerPool pool;                     // flattening threads, started on first use
  std::vector<ResizeTabDev> resize_tabs;
  // tracker
  hso_track_params tprm;
  int tB = 0, t_trace_cap = 0;
  int t_cluster = 0, t_threads = 0;  // 0 = auto
  int t_shape[kMaxLevels][2] = {{0}};  // per-level override {cluster, threads}
  int t_used[kMaxLevels][5] = {{0}};   // launch shape of the last run per level: {cluster, threads, mode, absres_smem, hist_bits}
  DevBuf t_arena, t_jobs_dev, t_T0, t_a0, t_out_dev;
  PinBuf t_stage_host, t_jobs_host, t_out_host;
  std::vector<size_t> t_trace_off;  // byte offset of each job's trace in the arena
  std::vector<size_t> t_goff;       // staging plan of the batch in flight (track_plan): byte offset of each job's geometry block
  size_t t_geo_bytes = 0;
  // direct-input mode (pinned caller arrays are copied as they are and flattened on the device): -1 auto, 0 never, 1 always
  int t_no_ring1 = getenv("HSO_TRACK_NO_RING1") ? 1 : 0;          // tuning: never the single-buffered ring (inverse-compositional level 1)
  int t_no_pair = getenv("HSO_TRACK_NO_PAIR") ? 1 : 0;            // tuning: never two 256-thread CTAs per SM in forward mode
  int t_force_stream = getenv("HSO_TRACK_FORCE_STREAM") ? 1 : 0;  // tuning / tests: mode 3 at every forward level where it fits
  int t_no_stream = getenv("HSO_TRACK_NO_STREAM") ? 1 : 0;        // tuning: never use the streamed-cache mode (mode 3) of the forward tracker
  int t_abs_global = getenv("HSO_TRACK_ABSRES_GLOBAL") ? 1 : 0;  // tuning: keep the |r| scratch of the threshold selection in global memory
  int t_direct_mode = -1;
  bool t_direct = false;           // decision for the batch in flight
  bool t_compact = false;          // the batch in flight uses hso_track_job::xyz / px32 (t_raw = [xyz 3 sumF doubles | px32 2 sumF floats])
  DevBuf t_raw;                    // [px 2 sumF | f 3 sumF | dist sumF] doubles
  std::vector<size_t> t_roff;      // features before job b in the raw arrays
  size_t t_sumF = 0;
  // chunk pipeline of hso_add_frames_track_batch: H2D on its own stream, one event per chunk
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  std::vector<cudaStream_t> pipe_streams;  // extra compute streams: chunk c runs on stream c % S so that its tail overlaps the next chunk
  std::vector<cudaEvent_t> pipe_ev;
  int pipe_chunk = 0, pipe_n_streams = 0;  // 0 = default
  size_t t_arena_bytes = 0;
  int t_maxF = 0;
  int t_profile = 0, t_prof_pending = 0, t_no_dual = 0;
  cudaEvent_t t_ev[kMaxLevels + 1] = {nullptr};
  double t_level_ms[kMaxLevels] = {0};
  uint64_t t_level_launches[kMaxLevels] = {0};
  // FAST scratch
  DevBuf f_score, f_rowbuf, f_rowcount, f_out, f_total;
  PinBuf f_out_host;
  // input side (row N4): raw staging, undistortion maps (built on first use), resize tables of the last raw size
  DevBuf in_raw, in_mid, in_ptrs, u_map1, u_map2, in_tab_blob;
  PinBuf in_ptrs_host;
  bool u_maps_ready = false;
  std::vector<short> u_map1_host;
  std::vector<uint16_t> u_map2_host;
  ResizeTabDev in_tab{};
  int in_tab_w = 0, in_tab_h = 0;
  // depth filter (row N3) scratch
  DevBuf d_arena;
  PinBuf d_stage_host, d_out_host;
  // reprojection (row N1) scratch
  DevBuf r_arena;
  PinBuf r_stage_host, r_out_host;
  // align / pose scratch
  DevBuf a_jobs_dev, a_out_dev, p_arena, p_jobs_dev, p_out_dev;
  PinBuf a_jobs_host, a_out_host, p_stage_host, p_out_host;
  // stage timers
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;  // ev2 / ev3: the alignment kernel nested inside a reprojection call
  double stage_ms[kStages] = {0};
  uint64_t stage_calls[kStages] = {0};
};

namespace {

int fail(hso_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
  }
  return code;
}
#define CU(call)                                                              \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) return fail(ctx, HSO_ERR_CUDA, #call, e__);       \
  } while (0)

inline int cv_round(double v) { return (int)std::nearbyint(v); }
inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

// frame_utils::createImgPyramid level sizes (src/frame.cpp:296-314).
void build_geom(const hso_cam& cam, const hso_cfg& cfg, PyrGeom& g) {
  memset(&g, 0, sizeof g);
  g.n_levels = std::max(cfg.n_pyr_levels, cfg.klt_max_level + 1);
  if (g.n_levels > kMaxLevels) g.n_levels = kMaxLevels;
  const int W = cam.width, H = cam.height;
  g.half_path = (W % 16 == 0) && (H % 16 == 0);
  g.w[0] = W; g.h[0] = H;
  for (int i = 1; i < g.n_levels; ++i) {
    if (g.half_path) {
      g.w[i] = g.w[i - 1] / 2; g.h[i] = g.h[i - 1] / 2;
      g.sse_rounding[i] = (g.w[i - 1] % 16) == 0;
    } else {
      const float scale = (float)(1.0 / (1 << i));
      g.w[i] = cv_round((float)W * scale);
      g.h[i] = cv_round((float)H * scale);
    }
  }
  size_t off = 0;
  for (int i = 0; i < g.n_levels; ++i) {
    g.off[i] = off;
    const size_t bytes = (size_t)g.w[i] * g.h[i] + (size_t)kPadRows * g.w[i] + 16;
    g.stage_bytes[i] = (uint32_t)((bytes + 15) / 16 * 16);
    off += (g.stage_bytes[i] + 127) / 128 * 128;
  }
  g.bytes = off;
}

// Coefficient tables of cv::resize(INTER_LINEAR, 8UC1) for one level (OpenCV resize.cpp; fixed point, 11 bits).
int build_resize_tab(hso_ctx* ctx, int sw, int sh, int dw, int dh, std::vector<char>& blob, ResizeTabDev& tab_offsets) {
  const double inv_fx = (double)sw / dw, inv_fy = (double)sh / dh;
  const int isx = (int)std::lrint(inv_fx), isy = (int)std::lrint(inv_fy);
  const bool area = std::fabs(inv_fx - isx) < 2.220446049250313e-16 && std::fabs(inv_fy - isy) < 2.220446049250313e-16;
  tab_offsets.area_fast = (area && isx == 2 && isy == 2) ? 1 : 0;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> ialpha(2 * dw), ibeta(2 * dh);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * inv_fx - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    ialpha[2 * dx] = sat_short(cv_round((1.f - fx) * 2048));
    ialpha[2 * dx + 1] = sat_short(cv_round(fx * 2048));
  }
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * inv_fy - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    ibeta[2 * dy] = sat_short(cv_round((1.f - fy) * 2048));
    ibeta[2 * dy + 1] = sat_short(cv_round(fy * 2048));
  }
  auto put = [&](const void* p, size_t n) {
    size_t o = (blob.size() + 15) / 16 * 16;
    blob.resize(o + n);
    memcpy(blob.data() + o, p, n);
    return o;
  };
  tab_offsets.xofs = (const int*)put(xofs.data(), xofs.size() * 4);
  tab_offsets.ialpha = (const short*)put(ialpha.data(), ialpha.size() * 2);
  tab_offsets.yofs = (const int*)put(yofs.data(), yofs.size() * 4);
  tab_offsets.ibeta = (const short*)put(ibeta.data(), ibeta.size() * 2);
  (void)ctx;
  return 0;
}

void unuse_frame(hso_ctx* ctx, hso_frame_id id) {
  ctx->frames[id].used = false;
  if ((size_t)id < ctx->free_lo) ctx->free_lo = (size_t)id;
}

int alloc_frame(hso_ctx* ctx, hso_frame_id* out) {
  for (size_t i = ctx->free_lo; i < ctx->frames.size(); ++i) {
    FrameSlot& s = ctx->frames[i];
    if (s.used) continue;
    if (!s.pyr) {
      // slots live in one arena (zeroed at creation: the pad rows behind every level stay zero), so a batch of equally spaced
      // host images lands in consecutive slots with ONE 2-D copy
      s.pyr = (uint8_t*)ctx->pyr_arena.p + i * ctx->pyr_slot_bytes;
      s.sums = (double*)ctx->sums_arena.p + (size_t)2 * ctx->n_tiles * i;
      s.stats = (float*)ctx->stats_table.p + 2 * i;
      if (ctx->cfg.materialize_sobel) CU(cudaMalloc((void**)&s.sobel, sizeof(int16_t) * ctx->sobel_elems));
    }
    s.used = true;
    ctx->free_lo = i + 1;
    *out = (hso_frame_id)i;
    return HSO_OK;
  }
  return fail(ctx, HSO_ERR_CAPACITY, "frame table full (hso_cfg.max_frames)");
}

FrameSlot* get_frame(hso_ctx* ctx, hso_frame_id id) {
  if (id < 0 || (size_t)id >= ctx->frames.size() || !ctx->frames[id].used) return nullptr;
  return &ctx->frames[id];
}

// slot0: first record of the (pre-reserved) job / counter arrays this call may use — chunks of a pipelined batch keep their records
// apart because the previous chunk's may still be in flight. The job records travel on `copy`, the kernels run on ctx->stream.
int run_pyramid(hso_ctx* ctx, int B, const hso_frame_id* ids, const uint8_t* const* srcs, int src_stride, int aligned, int slot0 = 0,
                cudaStream_t copy = nullptr) {
  const size_t need = (size_t)slot0 + B;
  CU(ctx->pyr_jobs_dev.reserve(sizeof(PyrJobDev) * need));
  if (ctx->pyr_counters.cap < sizeof(unsigned) * need) {
    CU(ctx->pyr_counters.reserve(sizeof(unsigned) * need));
    CU(cudaMemsetAsync(ctx->pyr_counters.p, 0, ctx->pyr_counters.cap, ctx->stream));
  }
  PyrJobDev* jobs;
  if (copy) {
    // pipelined batch: the caller reserved pyr_jobs_host for the whole batch and every chunk has its own records
    if (ctx->pyr_jobs_host.cap < sizeof(PyrJobDev) * need) return fail(ctx, HSO_ERR_INVALID, "pipelined pyramid records not reserved");
    jobs = (PyrJobDev*)ctx->pyr_jobs_host.p + slot0;
  } else {
    // asynchronous entry points (hso_frame_[re]build_batch_device return unsynchronised): the previous call's H2D may not have read its
    // records yet, so each call writes into its own region of a small ring
    void* region = nullptr;
    CU(ctx->pyr_jobs_ring.acquire(sizeof(PyrJobDev) * B, &region));
    jobs = (PyrJobDev*)region;
  }
  for (int i = 0; i < B; ++i) {
    FrameSlot* s = get_frame(ctx, ids[i]);
    jobs[i].src = srcs[i];
    jobs[i].pyr = s->pyr;
    jobs[i].sobel = s->sobel;
    jobs[i].sums = s->sums;
    jobs[i].stats = s->stats;
  }
  PyrJobDev* jobs_dev = (PyrJobDev*)ctx->pyr_jobs_dev.p + slot0;
  CU(cudaMemcpyAsync(jobs_dev, jobs, sizeof(PyrJobDev) * B, cudaMemcpyHostToDevice, copy ? copy : ctx->stream));
  if (copy) return HSO_OK;  // the caller launches after its event wait (launch_pyramid_slots)
  CU(ctx->pyr_jobs_ring.commit(ctx->stream));
  CU(launch_pyramid(ctx->geom, jobs_dev, B, src_stride, ctx->resize_tabs.data() /* host array; passed by value */,
                    ctx->cfg.materialize_sobel, (unsigned*)ctx->pyr_counters.p + slot0, aligned, ctx->stream, &ctx->launches));
  return HSO_OK;
}

int read_stats(hso_ctx* ctx, int B, const hso_frame_id* ids, float* integral, float* grad_mean) {
  if (!integral && !grad_mean) return HSO_OK;
  // one copy of the id range instead of one tiny copy per frame
  int lo = ids[0], hi = ids[0];
  for (int i = 1; i < B; ++i) { lo = std::min(lo, (int)ids[i]); hi = std::max(hi, (int)ids[i]); }
  CU(ctx->stats_host.reserve(sizeof(float) * 2 * (hi - lo + 1)));
  float* h = (float*)ctx->stats_host.p;
  CU(cudaMemcpyAsync(h, (float*)ctx->stats_table.p + 2 * lo, sizeof(float) * 2 * (hi - lo + 1), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < B; ++i) {
    if (integral) integral[i] = h[2 * (ids[i] - lo)];
    if (grad_mean) grad_mean[i] = h[2 * (ids[i] - lo) + 1];
  }
  return HSO_OK;
}

struct StageTimer {
  hso_ctx* c; int stage;
  StageTimer(hso_ctx* ctx, int s) : c(ctx), stage(s) { cudaEventRecord(c->ev0, c->stream); }
  void stop_after_sync() {
    cudaEventRecord(c->ev1, c->stream);
    cudaEventSynchronize(c->ev1);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) { c->stage_ms[stage] += ms; c->stage_calls[stage]++; }
    if (nested && cudaEventElapsedTime(&ms, c->ev2, c->ev3) == cudaSuccess) { c->stage_ms[ST_FEATURE_ALIGN] += ms; c->stage_calls[ST_FEATURE_ALIGN]++; }
  }
  // "feature_align" runs inside "reproject" in the reference (src/reprojector.cpp:259-330): bracket the alignment kernel of a reprojection call
  bool nested = false;
  void align_begin() { cudaEventRecord(c->ev2, c->stream); }
  void align_end() { cudaEventRecord(c->ev3, c->stream); nested = true; }
};

}  // namespace

extern "C" {

void hso_cfg_default(hso_cfg* cfg) {
  memset(cfg, 0, sizeof *cfg);
  cfg->n_pyr_levels = 3;   // src/config.cpp:32
  cfg->klt_max_level = 4;  // src/config.cpp:40
  cfg->max_frames = 16;
  cfg->max_features = 8192;
  cfg->materialize_sobel = 0;
}

int hso_create(int device, const hso_cam* cam, const hso_cfg* cfg_in, hso_ctx** out) {
  if (!cam || !out) return HSO_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return HSO_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return HSO_ERR_NO_DEVICE;
  if (prop.major != 10) return HSO_ERR_NO_DEVICE;  // the kernels are sm_100a only
  if (cam->width <= 32 || cam->height <= 32) return HSO_ERR_INVALID;
  hso_ctx* ctx = new hso_ctx();
  ctx->device = device;
  ctx->cam = *cam;
  if (cfg_in) ctx->cfg = *cfg_in; else hso_cfg_default(&ctx->cfg);
  if (ctx->cfg.max_frames <= 0) ctx->cfg.max_frames = 16;
  if (ctx->cfg.max_features <= 0) ctx->cfg.max_features = 8192;
  if (ctx->cfg.n_pyr_levels <= 0) ctx->cfg.n_pyr_levels = 3;
  if (ctx->cfg.klt_max_level <= 0) ctx->cfg.klt_max_level = 4;
  CamDev& cd = ctx->camdev;
  cd.model = cam->model; cd.width = cam->width; cd.height = cam->height; cd.undistort = cam->undistort;
  cd.fx = cam->fx; cd.fy = cam->fy; cd.cx = cam->cx; cd.cy = cam->cy;
  for (int i = 0; i < 5; ++i) cd.d[i] = cam->d[i];
  cd.distortion = (cam->model == 0 && std::fabs(cam->d[0]) > 0.0000001) ? 1 : 0;  // src/camera.cpp:36
  build_geom(ctx->cam, ctx->cfg, ctx->geom);
  ctx->n_tiles = pyramid_tiles(ctx->geom);
  ctx->sobel_elems = 0;
  for (int l = 0; l < 3 && l < ctx->geom.n_levels; ++l) ctx->sobel_elems += (size_t)2 * ctx->geom.w[l] * ctx->geom.h[l];
  ctx->frames.resize(ctx->cfg.max_frames);
  auto bail = [&](const char* what) { fprintf(stderr, "hso_create: %s\n", what); hso_destroy(ctx); return HSO_ERR_CUDA; };
  if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice");
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate");
  ctx->stream = ctx->own_stream;
  if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess || cudaEventCreate(&ctx->ev2) != cudaSuccess ||
      cudaEventCreate(&ctx->ev3) != cudaSuccess)
    return bail("cudaEventCreate");
  if (ctx->stats_table.reserve(sizeof(float) * 2 * ctx->cfg.max_frames) != cudaSuccess) return bail("cudaMalloc stats table");
  ctx->pyr_slot_bytes = (ctx->geom.bytes + 255) / 256 * 256;
  if (ctx->pyr_arena.reserve(ctx->pyr_slot_bytes * ctx->cfg.max_frames) != cudaSuccess) return bail("cudaMalloc pyramid arena (hso_cfg.max_frames too large?)");
  if (cudaMemset(ctx->pyr_arena.p, 0, ctx->pyr_arena.cap) != cudaSuccess) return bail("cudaMemset pyramid arena");
  if (ctx->sums_arena.reserve(sizeof(double) * 2 * ctx->n_tiles * ctx->cfg.max_frames) != cudaSuccess) return bail("cudaMalloc sums arena");
  if (!ctx->geom.half_path) {
    std::vector<char> blob;
    ctx->resize_tabs.assign(ctx->geom.n_levels, ResizeTabDev{});
    for (int l = 1; l < ctx->geom.n_levels; ++l)
      build_resize_tab(ctx, ctx->geom.w[l - 1], ctx->geom.h[l - 1], ctx->geom.w[l], ctx->geom.h[l], blob, ctx->resize_tabs[l]);
    const size_t tab_bytes = sizeof(ResizeTabDev) * ctx->geom.n_levels;
    const size_t blob_off = (tab_bytes + 255) / 256 * 256;
    if (ctx->resize_tab_dev.reserve(blob_off + blob.size()) != cudaSuccess) return bail("cudaMalloc resize tables");
    char* base = (char*)ctx->resize_tab_dev.p + blob_off;
    for (int l = 1; l < ctx->geom.n_levels; ++l) {
      ResizeTabDev& t = ctx->resize_tabs[l];
      t.xofs = (const int*)(base + (size_t)t.xofs);
      t.ialpha = (const short*)(base + (size_t)t.ialpha);
      t.yofs = (const int*)(base + (size_t)t.yofs);
      t.ibeta = (const short*)(base + (size_t)t.ibeta);
    }
    if (cudaMemcpy(base, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) return bail("cudaMemcpy resize tables");
  }
  ctx->tprm.inverse_comp = 0; ctx->tprm.max_level = 4; ctx->tprm.min_level = 1; ctx->tprm.n_iter = 50;
  *out = ctx;
  return HSO_OK;
}

void hso_destroy(hso_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  for (FrameSlot& s : ctx->frames) {
    if (s.sobel) cudaFree(s.sobel);
  }
  DevBuf* db[] = {&ctx->f_score, &ctx->f_rowbuf, &ctx->f_rowcount, &ctx->f_out, &ctx->f_total, &ctx->pyr_arena, &ctx->sums_arena, &ctx->stats_table, &ctx->pyr_jobs_dev, &ctx->pyr_counters, &ctx->resize_tab_dev, &ctx->t_arena, &ctx->t_jobs_dev, &ctx->t_T0, &ctx->t_a0,
                  &ctx->t_out_dev, &ctx->t_raw, &ctx->a_jobs_dev, &ctx->a_out_dev, &ctx->in_raw, &ctx->in_mid, &ctx->in_ptrs, &ctx->u_map1, &ctx->u_map2, &ctx->in_tab_blob, &ctx->d_arena, &ctx->r_arena, &ctx->p_arena, &ctx->p_jobs_dev, &ctx->p_out_dev};
  for (DevBuf* b : db) b->release();
  PinBuf* pb[] = {&ctx->f_out_host, &ctx->pyr_jobs_host, &ctx->stats_host, &ctx->t_stage_host, &ctx->t_jobs_host, &ctx->t_out_host,
                  &ctx->a_jobs_host, &ctx->a_out_host, &ctx->in_ptrs_host, &ctx->d_stage_host, &ctx->d_out_host, &ctx->r_stage_host, &ctx->r_out_host, &ctx->p_stage_host, &ctx->p_out_host};
  for (PinBuf* b : pb) b->release();
  ctx->pyr_jobs_ring.release(); ctx->t_jobs_ring.release();
  for (cudaEvent_t e : ctx->t_ev) if (e) cudaEventDestroy(e);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev2) cudaEventDestroy(ctx->ev2);
  if (ctx->ev3) cudaEventDestroy(ctx->ev3);
  for (cudaEvent_t e : ctx->chunk_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->pipe_ev) cudaEventDestroy(e);
  for (cudaStream_t st : ctx->pipe_streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* hso_last_error(const hso_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int hso_set_stream(hso_ctx* ctx, void* s) {
  if (!ctx) return HSO_ERR_INVALID;
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return HSO_OK;
}
void* hso_get_stream(hso_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int hso_synchronize(hso_ctx* ctx) {
  if (!ctx) return HSO_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return HSO_OK;
}
uint64_t hso_kernel_launches(const hso_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- F1 ---------------------------------------------------------------------------------------------------------------------
int hso_frame_upload_batch(hso_ctx* ctx, int B, const uint8_t* const* imgs, int W, int H, int stride, hso_frame_id* out, float* integral,
                           float* grad_mean) {
  if (!ctx || B <= 0 || !imgs || !out) return HSO_ERR_INVALID;
  // Frame::initFrame: "image must be CV_8UC1 of the camera's size" else throws (src/frame.cpp:85-86)
  if (W != ctx->cam.width || H != ctx->cam.height || stride < W) return fail(ctx, HSO_ERR_INVALID, "image size does not match the camera model");
  CU(cudaSetDevice(ctx->device));
  std::vector<const uint8_t*> srcs(B);
  for (int i = 0; i < B; ++i) {
    if (!imgs[i]) return fail(ctx, HSO_ERR_INVALID, "null image");
    int rc = alloc_frame(ctx, &out[i]);
    if (rc != HSO_OK) { for (int j = 0; j < i; ++j) unuse_frame(ctx, out[j]); return rc; }
  }
  StageTimer tm(ctx, 0);
  // straight into the level-0 slot of each pyramid (row stride == W); the kernel then builds in place. Tightly packed,
  // equally spaced host images going to consecutive slots take ONE 2-D copy (row = one image, dpitch = slot stride).
  bool one_copy = B > 1 && stride == W;
  const ptrdiff_t spacing = B > 1 ? imgs[1] - imgs[0] : 0;
  for (int i = 1; i < B && one_copy; ++i) one_copy = (imgs[i] - imgs[i - 1] == spacing) && (out[i] == out[i - 1] + 1);
  one_copy = one_copy && spacing >= (ptrdiff_t)W * H;
  if (one_copy) {
    CU(cudaMemcpy2DAsync(get_frame(ctx, out[0])->pyr + ctx->geom.off[0], ctx->pyr_slot_bytes, imgs[0], (size_t)spacing, (size_t)W * H, B,
                         cudaMemcpyHostToDevice, ctx->stream));
  }
  for (int i = 0; i < B; ++i) {
    FrameSlot* s = get_frame(ctx, out[i]);
    if (!one_copy) CU(cudaMemcpy2DAsync(s->pyr + ctx->geom.off[0], W, imgs[i], stride, W, H, cudaMemcpyHostToDevice, ctx->stream));
    srcs[i] = s->pyr + ctx->geom.off[0];
  }
  int rc = run_pyramid(ctx, B, out, srcs.data(), W, 1);
  if (rc != HSO_OK) return rc;
  rc = read_stats(ctx, B, out, integral, grad_mean);
  if (rc != HSO_OK) return rc;
  tm.stop_after_sync();
  return HSO_OK;
}

// ---- N4: raw image -> (resize) -> (undistort) -> Frame --------------------------------------------------------------------------
static int ensure_undistort_maps(hso_ctx* ctx) {
  if (ctx->u_maps_ready) return HSO_OK;
  build_undistort_maps(ctx->cam, ctx->u_map1_host, ctx->u_map2_host);
  CU(ctx->u_map1.reserve(ctx->u_map1_host.size() * sizeof(short)));
  CU(ctx->u_map2.reserve(ctx->u_map2_host.size() * sizeof(uint16_t)));
  CU(cudaMemcpy(ctx->u_map1.p, ctx->u_map1_host.data(), ctx->u_map1_host.size() * sizeof(short), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(ctx->u_map2.p, ctx->u_map2_host.data(), ctx->u_map2_host.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  ctx->u_maps_ready = true;
  return HSO_OK;
}

int hso_undistort_maps(hso_ctx* ctx, int16_t* map1, uint16_t* map2) {
  if (!ctx) return HSO_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  int rc = ensure_undistort_maps(ctx);
  if (rc != HSO_OK) return rc;
  if (map1) memcpy(map1, ctx->u_map1_host.data(), ctx->u_map1_host.size() * sizeof(short));
  if (map2) memcpy(map2, ctx->u_map2_host.data(), ctx->u_map2_host.size() * sizeof(uint16_t));
  return HSO_OK;
}

int hso_frame_upload_raw_batch(hso_ctx* ctx, int B, const uint8_t* const* imgs, int raw_w, int raw_h, int stride, int undistort, hso_frame_id* out,
                               float* integral, float* grad_mean) {
  if (!ctx || B <= 0 || !imgs || !out || raw_w <= 1 || raw_h <= 1 || stride < raw_w) return HSO_ERR_INVALID;
  const int W = ctx->cam.width, H = ctx->cam.height;
  const bool need_resize = raw_w != W || raw_h != H;
  if (!need_resize && !undistort) return hso_frame_upload_batch(ctx, B, imgs, W, H, stride, out, integral, grad_mean);
  CU(cudaSetDevice(ctx->device));
  for (int i = 0; i < B; ++i)
    if (!imgs[i]) return fail(ctx, HSO_ERR_INVALID, "null image");
  if (undistort) { int rc = ensure_undistort_maps(ctx); if (rc != HSO_OK) return rc; }
  CU(cudaStreamSynchronize(ctx->stream));  // staging buffers below are reused between calls
  if (need_resize && (ctx->in_tab_w != raw_w || ctx->in_tab_h != raw_h)) {
    std::vector<char> blob;
    ResizeTabDev t{};
    build_resize_tab(ctx, raw_w, raw_h, W, H, blob, t);
    CU(ctx->in_tab_blob.reserve(blob.size()));
    CU(cudaMemcpy(ctx->in_tab_blob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    char* base = (char*)ctx->in_tab_blob.p;
    t.xofs = (const int*)(base + (size_t)t.xofs); t.ialpha = (const short*)(base + (size_t)t.ialpha);
    t.yofs = (const int*)(base + (size_t)t.yofs); t.ibeta = (const short*)(base + (size_t)t.ibeta);
    ctx->in_tab = t; ctx->in_tab_w = raw_w; ctx->in_tab_h = raw_h;
  }
  for (int i = 0; i < B; ++i) {
    int rc = alloc_frame(ctx, &out[i]);
    if (rc != HSO_OK) { for (int j = 0; j < i; ++j) unuse_frame(ctx, out[j]); return rc; }
  }
  const size_t raw_bytes = ((size_t)raw_w * raw_h + 255) / 256 * 256, mid_bytes = ((size_t)W * H + 255) / 256 * 256;
  CU(ctx->in_raw.reserve(raw_bytes * B));
  if (need_resize && undistort) CU(ctx->in_mid.reserve(mid_bytes * B));
  CU(ctx->in_ptrs.reserve(sizeof(void*) * 3 * B));
  CU(ctx->in_ptrs_host.reserve(sizeof(void*) * 3 * B));
  const uint8_t** hp = (const uint8_t**)ctx->in_ptrs_host.p;  // [raw | mid | level-0 slot] x B
  std::vector<const uint8_t*> srcs(B);
  for (int i = 0; i < B; ++i) {
    hp[i] = (const uint8_t*)ctx->in_raw.p + raw_bytes * i;
    hp[B + i] = (need_resize && undistort) ? (const uint8_t*)ctx->in_mid.p + mid_bytes * i : nullptr;
    hp[2 * B + i] = get_frame(ctx, out[i])->pyr + ctx->geom.off[0];
    srcs[i] = hp[2 * B + i];
  }
  StageTimer tm(ctx, 0);
  CU(cudaMemcpyAsync(ctx->in_ptrs.p, hp, sizeof(void*) * 3 * B, cudaMemcpyHostToDevice, ctx->stream));
  for (int i = 0; i < B; ++i)
    CU(cudaMemcpy2DAsync((void*)hp[i], raw_w, imgs[i], stride, raw_w, raw_h, cudaMemcpyHostToDevice, ctx->stream));
  const uint8_t* const* d_raw = (const uint8_t* const*)ctx->in_ptrs.p;
  uint8_t* const* d_mid = (uint8_t* const*)ctx->in_ptrs.p + B;
  uint8_t* const* d_dst = (uint8_t* const*)ctx->in_ptrs.p + 2 * B;
  if (need_resize)  // ImageReader::readImage: cv::resize(image, image, m_img_new_size)
    CU(launch_resize(d_raw, raw_w, raw_h, raw_w, ctx->in_tab, undistort ? d_mid : d_dst, W, H, B, ctx->stream, &ctx->launches));
  if (undistort)    // cam_->undistortImage(image, image)
    CU(launch_remap(need_resize ? (const uint8_t* const*)d_mid : d_raw, W, H, W, (const short2*)ctx->u_map1.p, (const uint16_t*)ctx->u_map2.p, d_dst,
                    W, H, B, ctx->stream, &ctx->launches));
  int rc = run_pyramid(ctx, B, out, srcs.data(), W, 1);
  if (rc != HSO_OK) return rc;
  rc = read_stats(ctx, B, out, integral, grad_mean);
  if (rc != HSO_OK) return rc;
  if (!integral && !grad_mean) CU(cudaStreamSynchronize(ctx->stream));
  tm.stop_after_sync();
  return HSO_OK;
}

int hso_frame_upload(hso_ctx* ctx, const uint8_t* img, int W, int H, int stride, hso_frame_id* out, float* integral, float* grad_mean) {
  return hso_frame_upload_batch(ctx, 1, &img, W, H, stride, out, integral, grad_mean);
}

int hso_frame_build_batch_device(hso_ctx* ctx, int B, const void* const* dev_imgs, int W, int H, int stride, hso_frame_id* out) {
  if (!ctx || B <= 0 || !dev_imgs || !out) return HSO_ERR_INVALID;
  if (W != ctx->cam.width || H != ctx->cam.height || stride < W) return fail(ctx, HSO_ERR_INVALID, "image size does not match the camera model");
  CU(cudaSetDevice(ctx->device));
  int aligned = 1;
  for (int i = 0; i < B; ++i) {
    if (((uintptr_t)dev_imgs[i] & 15) != 0) aligned = 0;
    int rc = alloc_frame(ctx, &out[i]);
    if (rc != HSO_OK) { for (int j = 0; j < i; ++j) unuse_frame(ctx, out[j]); return rc; }
  }
  return run_pyramid(ctx, B, out, (const uint8_t* const*)dev_imgs, stride, aligned);
}

// Rebuild existing frames from device-resident images (benchmark loop: no allocation, asynchronous).
int hso_frame_rebuild_batch_device(hso_ctx* ctx, int B, const void* const* dev_imgs, int W, int H, int stride, const hso_frame_id* ids) {
  if (!ctx || B <= 0 || !dev_imgs || !ids) return HSO_ERR_INVALID;
  if (W != ctx->cam.width || H != ctx->cam.height || stride < W) return fail(ctx, HSO_ERR_INVALID, "image size does not match the camera model");
  CU(cudaSetDevice(ctx->device));
  int aligned = 1;
  for (int i = 0; i < B; ++i) {
    if (!get_frame(ctx, ids[i])) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown frame id");
    if (((uintptr_t)dev_imgs[i] & 15) != 0) aligned = 0;
  }
  return run_pyramid(ctx, B, ids, (const uint8_t* const*)dev_imgs, stride, aligned);
}

int hso_frame_stats(hso_ctx* ctx, hso_frame_id id, float* integral, float* grad_mean) {
  if (!ctx) return HSO_ERR_INVALID;
  if (!get_frame(ctx, id)) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown frame id");
  CU(cudaSetDevice(ctx->device));
  return read_stats(ctx, 1, &id, integral, grad_mean);
}

int hso_frame_level_size(hso_ctx* ctx, hso_frame_id id, int level, int* w, int* h) {
  if (!ctx || level < 0 || level >= ctx->geom.n_levels) return HSO_ERR_INVALID;
  (void)id;
  if (w) *w = ctx->geom.w[level];
  if (h) *h = ctx->geom.h[level];
  return HSO_OK;
}

int hso_frame_download_level(hso_ctx* ctx, hso_frame_id id, int level, uint8_t* dst) {
  if (!ctx || !dst || level < 0 || level >= ctx->geom.n_levels) return HSO_ERR_INVALID;
  FrameSlot* s = get_frame(ctx, id);
  if (!s) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown frame id");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(dst, s->pyr + ctx->geom.off[level], (size_t)ctx->geom.w[level] * ctx->geom.h[level], cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HSO_OK;
}

int hso_frame_download_sobel(hso_ctx* ctx, hso_frame_id id, int level, int16_t* gx, int16_t* gy) {
  if (!ctx || level < 0 || level >= 3 || level >= ctx->geom.n_levels) return HSO_ERR_INVALID;
  FrameSlot* s = get_frame(ctx, id);
  if (!s) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown frame id");
  if (!s->sobel) return fail(ctx, HSO_ERR_INVALID, "context was created without materialize_sobel");
  CU(cudaSetDevice(ctx->device));
  size_t so = 0;
  for (int l = 0; l < level; ++l) so += (size_t)2 * ctx->geom.w[l] * ctx->geom.h[l];
  const size_t n = (size_t)ctx->geom.w[level] * ctx->geom.h[level];
  if (gx) CU(cudaMemcpyAsync(gx, s->sobel + so, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
  if (gy) CU(cudaMemcpyAsync(gy, s->sobel + so + n, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HSO_OK;
}

int hso_frame_release(hso_ctx* ctx, hso_frame_id id) {
  if (!ctx) return HSO_ERR_INVALID;
  FrameSlot* s = get_frame(ctx, id);
  if (!s) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown frame id");
  unuse_frame(ctx, id);  // buffers are kept for reuse; stream order protects in-flight work on this context
  return HSO_OK;
}

// ---- F2 ---------------------------------------------------------------------------------------------------------------------
int hso_frame_release_batch(hso_ctx* ctx, int n, const hso_frame_id* ids) {
  if (!ctx || n < 0 || (n > 0 && !ids)) return HSO_ERR_INVALID;
  int rc = HSO_OK;
  for (int i = 0; i < n; ++i) {
    const int r = hso_frame_release(ctx, ids[i]);
    if (r != HSO_OK) rc = r;
  }
  return rc;
}

int hso_track_set_cluster(hso_ctx* ctx, int ctas, int threads) {
  if (!ctx) return HSO_ERR_INVALID;
  if (!(ctas == 0 || ctas == 1 || ctas == 2 || ctas == 4 || ctas == 8)) return fail(ctx, HSO_ERR_INVALID, "cluster size must be 0,1,2,4,8");
  if (threads != 0 && (threads < 64 || threads > 512 || threads % 32)) return fail(ctx, HSO_ERR_INVALID, "threads must be a multiple of 32 in [64,512]");
  ctx->t_cluster = ctas;
  ctx->t_threads = threads;
  return HSO_OK;
}

int hso_set_pipeline(hso_ctx* ctx, int chunk, int streams) {
  if (!ctx || chunk < 0 || streams < 0 || streams > 8) return HSO_ERR_INVALID;
  ctx->pipe_chunk = chunk; ctx->pipe_n_streams = streams;
  return HSO_OK;
}
int hso_track_set_direct_inputs(hso_ctx* ctx, int mode) {
  if (!ctx || mode < -1 || mode > 1) return HSO_ERR_INVALID;
  ctx->t_direct_mode = mode;
  return HSO_OK;
}
int hso_track_set_ic_dual(hso_ctx* ctx, int enable) {
  if (!ctx) return HSO_ERR_INVALID;
  ctx->t_no_dual = enable ? 0 : 1;
  return HSO_OK;
}

int hso_track_set_stream_cache(hso_ctx* ctx, int mode) {
  if (!ctx || mode < -1 || mode > 1) return HSO_ERR_INVALID;
  ctx->t_no_stream = mode < 0 ? 1 : 0;
  ctx->t_force_stream = mode > 0 ? 1 : 0;
  return HSO_OK;
}

int hso_track_get_level_shape(hso_ctx* ctx, int level, int* ctas, int* threads, int* mode, int* absres_smem) {
  if (!ctx || level < 0 || level >= kMaxLevels) return HSO_ERR_INVALID;
  if (ctas) *ctas = ctx->t_used[level][0];
  if (threads) *threads = ctx->t_used[level][1];
  if (mode) *mode = ctx->t_used[level][2];
  if (absres_smem) *absres_smem = ctx->t_used[level][3];
  return HSO_OK;
}

int hso_track_set_level_shape(hso_ctx* ctx, int level, int ctas, int threads) {
  if (!ctx || level < 0 || level >= kMaxLevels) return HSO_ERR_INVALID;
  if (!(ctas == 0 || ctas == 1 || ctas == 2 || ctas == 4 || ctas == 8)) return fail(ctx, HSO_ERR_INVALID, "cluster size must be 0,1,2,4,8");
  if (threads != 0 && (threads < 64 || threads > 512 || threads % 32)) return fail(ctx, HSO_ERR_INVALID, "threads must be a multiple of 32 in [64,512]");
  ctx->t_shape[level][0] = ctas;
  ctx->t_shape[level][1] = threads;
  return HSO_OK;
}

// Staging plan of a batch: validation, compaction counts, arena layout, device job records (host copy). No copies are issued.
// Staging plan of a batch: validation, arena layout (by the upper bound n_features — the count of features with depth is only known
// after flattening), device job records (host copy). No copies are issued.
static int track_plan(hso_ctx* ctx, const hso_track_params* prm, int B, const hso_track_job* jobs, int trace_cap) {
  if (!ctx || !prm || B <= 0 || !jobs) return HSO_ERR_INVALID;
  if (prm->max_level >= ctx->geom.n_levels || prm->min_level < 0 || prm->min_level > prm->max_level || prm->max_level - prm->min_level > 5 ||
      prm->n_iter < 0)
    return fail(ctx, HSO_ERR_INVALID, "bad track params");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));  // the pinned staging below is reused between calls
  ctx->tprm = *prm;
  ctx->tB = B;
  ctx->t_trace_cap = trace_cap > 0 ? trace_cap : 0;
  size_t host_bytes = 0, arena = 0;
  int maxF = 0, n_layout = -1;  // feature layout of the batch: 0 wide (px / f / dist), 1 compact (xyz / px32)
  ctx->t_goff.assign(B, 0);
  for (int b = 0; b < B; ++b) {
    const hso_track_job& j = jobs[b];
    if (j.n_features < 0 || j.n_features > ctx->cfg.max_features) return fail(ctx, HSO_ERR_CAPACITY, "n_features exceeds hso_cfg.max_features");
    if (!get_frame(ctx, j.ref) || !get_frame(ctx, j.cur)) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown frame id in track job");
    const bool compact_b = j.xyz != nullptr && j.px32 != nullptr;
    if (j.n_features > 0 && !compact_b && (!j.px || !j.f || !j.dist)) return fail(ctx, HSO_ERR_INVALID, "null feature arrays");
    if (j.n_features > 0) {
      if (n_layout < 0) n_layout = compact_b ? 1 : 0;
      else if (n_layout != (compact_b ? 1 : 0)) return fail(ctx, HSO_ERR_INVALID, "jobs of a batch must use the same feature layout");
    }
    maxF = std::max(maxF, j.n_features);
    const int Fpad = std::max(32, (j.n_features + 31) / 32 * 32);
    ctx->t_goff[b] = host_bytes;
    host_bytes += sizeof(double) * 5 * Fpad;
    arena += sizeof(double) * 5 * Fpad;
  }
  ctx->t_maxF = maxF;
  ctx->t_geo_bytes = arena;
  // direct-input decision: pinned (or registered) caller arrays go to the device as they are, no host flattening pass
  ctx->t_direct = false;
  ctx->t_compact = n_layout == 1;
  if (ctx->t_compact) {
    ctx->t_direct = true;  // the compact layout always travels as it is (a pageable array is still copied correctly, only slower)
  } else if (ctx->t_direct_mode != 0) {
    bool pinned = ctx->t_direct_mode == 1;
    if (!pinned) {
      for (int b = 0; b < B && !pinned; ++b) {
        if (jobs[b].n_features == 0) continue;
        cudaPointerAttributes a0, a1, a2;
        pinned = cudaPointerGetAttributes(&a0, jobs[b].px) == cudaSuccess && a0.type == cudaMemoryTypeHost &&
                 cudaPointerGetAttributes(&a1, jobs[b].f) == cudaSuccess && a1.type == cudaMemoryTypeHost &&
                 cudaPointerGetAttributes(&a2, jobs[b].dist) == cudaSuccess && a2.type == cudaMemoryTypeHost;
        cudaGetLastError();  // an unregistered pointer is reported through the sticky-free error state on older drivers
        break;               // the first non-empty job decides for the batch (a pageable array in a later job is still copied correctly, only slower)
      }
    }
    ctx->t_direct = pinned;
  }
  ctx->t_roff.assign(B + 1, 0);
  for (int b = 0; b < B; ++b) ctx->t_roff[b + 1] = ctx->t_roff[b] + (size_t)jobs[b].n_features;
  ctx->t_sumF = ctx->t_roff[B];
  if (ctx->t_direct) CU(ctx->t_raw.reserve(sizeof(double) * 6 * std::max<size_t>(ctx->t_sumF, 1)));
  // per-job scratch behind the geometry block
  std::vector<size_t> off_cache(B), off_gx(B), off_gy(B), off_abs(B), off_vis(B), off_state(B), off_trace(B);
  for (int b = 0; b < B; ++b) {
    const int Fpad = std::max(32, (jobs[b].n_features + 31) / 32 * 32);
    auto take = [&](size_t bytes) { size_t o = (arena + 127) / 128 * 128; arena = o + bytes; return o; };
    off_cache[b] = take(sizeof(float) * kMaxPatternN * Fpad);
    if (prm->inverse_comp) { off_gx[b] = take(sizeof(float) * kMaxPatternN * Fpad); off_gy[b] = take(sizeof(float) * kMaxPatternN * Fpad); }
    // the streamed inverse-compositional cache (mode 3) treats the three planes as ONE group-major array of 3 N Fpad floats
    if (prm->inverse_comp && (off_gx[b] != off_cache[b] + sizeof(float) * kMaxPatternN * Fpad || off_gy[b] != off_gx[b] + sizeof(float) * kMaxPatternN * Fpad))
      return fail(ctx, HSO_ERR_INVALID, "tracker scratch planes are not contiguous");
    off_abs[b] = take(sizeof(float) * kMaxPatternN * Fpad);
    off_vis[b] = take(Fpad);
    off_state[b] = take(sizeof(TrackState));
    off_trace[b] = ctx->t_trace_cap ? take(sizeof(hso_trace) * ctx->t_trace_cap) : 0;
  }
  CU(ctx->t_arena.reserve(arena));
  ctx->t_arena_bytes = arena;
  ctx->t_trace_off = off_trace;
  CU(ctx->t_stage_host.reserve(host_bytes + sizeof(double) * 12 * B + sizeof(float) * B));
  CU(ctx->t_jobs_host.reserve(sizeof(TrackJobDev) * B));
  CU(ctx->t_jobs_dev.reserve(sizeof(TrackJobDev) * B));
  CU(ctx->t_T0.reserve(sizeof(double) * 12 * B));
  CU(ctx->t_a0.reserve(sizeof(float) * B));
  CU(ctx->t_out_dev.reserve(sizeof(hso_track_result) * B));
  CU(ctx->t_out_host.reserve(sizeof(hso_track_result) * B));
  char* dbase = (char*)ctx->t_arena.p;
  TrackJobDev* hj = (TrackJobDev*)ctx->t_jobs_host.p;
  for (int b = 0; b < B; ++b) {
    const hso_track_job& j = jobs[b];
    const int Fpad = std::max(32, (j.n_features + 31) / 32 * 32);
    TrackJobDev& d = hj[b];
    FrameSlot* fr = get_frame(ctx, j.ref);
    FrameSlot* fc = get_frame(ctx, j.cur);
    d.ref_pyr = fr->pyr;
    d.cur_pyr = fc->pyr;
    d.ref_stats = fr->stats;
    d.cur_stats = fc->stats;
    d.F = 0;  // set by track_stage_one
    d.Fpad = Fpad;
    d.px = (const double*)(dbase + ctx->t_goff[b]);
    d.xyz = d.px + 2 * Fpad;
    d.ref_cache = (float*)(dbase + off_cache[b]);
    d.ref_gx = prm->inverse_comp ? (float*)(dbase + off_gx[b]) : nullptr;
    d.ref_gy = prm->inverse_comp ? (float*)(dbase + off_gy[b]) : nullptr;
    d.absres = (float*)(dbase + off_abs[b]);
    d.vis = (uint8_t*)(dbase + off_vis[b]);
    d.state = (TrackState*)(dbase + off_state[b]);
    d.trace = ctx->t_trace_cap ? (hso_trace*)(dbase + off_trace[b]) : nullptr;
    d.raw_px = d.raw_f = d.raw_dist = nullptr; d.n_raw = 0; d.pad_ = 0;
    d.raw_xyz = nullptr; d.raw_px32 = nullptr;
    if (ctx->t_compact) {
      const double* rb = (const double*)ctx->t_raw.p;  // [xyz 3 sumF doubles | px32 2 sumF floats]
      d.raw_xyz = rb + 3 * ctx->t_roff[b];
      d.raw_px32 = (const float*)(rb + 3 * ctx->t_sumF) + 2 * ctx->t_roff[b];
      d.n_raw = j.n_features;
    } else if (ctx->t_direct) {
      const double* rb = (const double*)ctx->t_raw.p;
      d.raw_px = rb + 2 * ctx->t_roff[b];
      d.raw_f = rb + 2 * ctx->t_sumF + 3 * ctx->t_roff[b];
      d.raw_dist = rb + 5 * ctx->t_sumF + ctx->t_roff[b];
      d.n_raw = j.n_features;
    }
  }
  return HSO_OK;
}

// Non-temporal 8-byte store where the host ISA has one (x86-64); plain store elsewhere (e.g. aarch64 / Grace hosts).
static inline void nt_store(double* p, double v) {
#if defined(__x86_64__)
  long long bits;
  memcpy(&bits, &v, sizeof bits);
  _mm_stream_si64(reinterpret_cast<long long*>(p), bits);
#else
  *p = v;
#endif
}
static inline void nt_fence() {
#if defined(__x86_64__)
  _mm_sfence();
#endif
}

// Flatten the Feature list of job b to SoA in the pinned staging blob, keeping only features with a valid depth (dist >= 0): the
// reference skips the others in every stage (src/CoarseTracker.cpp:290,433,455,557), so dropping them only changes the summation
// order. Thread-safe across different b.
static void track_stage_one(hso_ctx* ctx, const hso_track_job* jobs, int b) {
  const int B = ctx->tB;
  char* hbase = (char*)ctx->t_stage_host.p;
  TrackJobDev* hj = (TrackJobDev*)ctx->t_jobs_host.p;
  const hso_track_job& j = jobs[b];
  const int Fpad = hj[b].Fpad;
  double* px = (double*)(hbase + ctx->t_goff[b]);
  double* xyz = px + 2 * Fpad;
  int k = 0;
  for (int i = 0; i < j.n_features; ++i) {
    const double d = j.dist[i];
    if (!(d >= 0)) continue;
    // non-temporal stores: the staging blob is written once and read by the DMA engine only; keeping it out of the caches (and skipping the
    // read-for-ownership) leaves more host memory bandwidth to the H2D copies that run concurrently with this loop
    nt_store(px + k, j.px[2 * i]); nt_store(px + Fpad + k, j.px[2 * i + 1]);
    // Vector3d xyz_ref((*it_ft)->f*dist)  (src/CoarseTracker.cpp:292)
    nt_store(xyz + k, j.f[3 * i] * d); nt_store(xyz + Fpad + k, j.f[3 * i + 1] * d); nt_store(xyz + 2 * Fpad + k, j.f[3 * i + 2] * d);
    ++k;
  }
  hj[b].F = k;
  nt_fence();
  for (; k < Fpad; ++k) { px[k] = px[Fpad + k] = 0; xyz[k] = xyz[Fpad + k] = 0; xyz[2 * Fpad + k] = 1; }
  double* T0 = (double*)(hbase + ctx->t_geo_bytes);
  float* a0 = (float*)(T0 + 12 * B);
  memcpy(T0 + 12 * b, j.T_cur_ref, sizeof(double) * 12);
  a0[b] = j.exposure_rat;
}

// direct-input mode: what track_stage_one records beside the geometry (initial pose, exposure ratio) for jobs [b0, b1)
static void track_fill_pose(hso_ctx* ctx, const hso_track_job* jobs, int b0, int b1) {
  const int B = ctx->tB;
  double* T0 = (double*)((char*)ctx->t_stage_host.p + ctx->t_geo_bytes);
  float* a0 = (float*)(T0 + 12 * B);
  for (int b = b0; b < b1; ++b) {
    memcpy(T0 + 12 * b, jobs[b].T_cur_ref, sizeof(double) * 12);
    a0[b] = jobs[b].exposure_rat;
  }
}

// Direct-input mode: the caller's feature arrays of jobs [b0, b1) (px / f / dist, or xyz / px32 in the compact layout) go to the device as they
// are. Arrays of consecutive jobs that are adjacent in host memory (a caller that keeps a batch in one blob) travel in one copy per run.
static int track_copy_raw(hso_ctx* ctx, const hso_track_job* jobs, int b0, int b1, cudaStream_t stream, bool fill_pose = true) {
  char* rb = (char*)ctx->t_raw.p;
  const size_t S = ctx->t_sumF;
  if (fill_pose) track_fill_pose(ctx, jobs, b0, b1);
  const int n_arr = ctx->t_compact ? 2 : 3;
  for (int arr = 0; arr < n_arr; ++arr) {
    // bytes per feature and device offset of the array: wide [px 16 | f 24 | dist 8], compact [xyz 24 | px32 8]
    const size_t w = ctx->t_compact ? (arr == 0 ? 24 : 8) : (arr == 0 ? 16 : (arr == 1 ? 24 : 8));
    char* dbase = rb + (ctx->t_compact ? (arr == 0 ? 0 : 24 * S) : (arr == 0 ? 0 : (arr == 1 ? 16 * S : 40 * S)));
    auto ptr = [&](int b) -> const char* {
      if (ctx->t_compact) return arr == 0 ? (const char*)jobs[b].xyz : (const char*)jobs[b].px32;
      return arr == 0 ? (const char*)jobs[b].px : (arr == 1 ? (const char*)jobs[b].f : (const char*)jobs[b].dist);
    };
    int run0 = b0;
    while (run0 < b1) {
      if (jobs[run0].n_features == 0) { ++run0; continue; }
      int run1 = run0 + 1;
      size_t feats = (size_t)jobs[run0].n_features;
      while (run1 < b1 && (jobs[run1].n_features == 0 || ptr(run1) == ptr(run0) + w * feats)) { feats += (size_t)jobs[run1].n_features; ++run1; }
      CU(cudaMemcpyAsync(dbase + w * ctx->t_roff[run0], ptr(run0), w * feats, cudaMemcpyHostToDevice, stream));
      run0 = run1;
    }
  }
  return HSO_OK;
}

// H2D of the staged records of jobs [b0, b1) on `stream`.
static int track_copy_range(hso_ctx* ctx, int b0, int b1, cudaStream_t stream) {
  const int B = ctx->tB, n = b1 - b0;
  char* hbase = (char*)ctx->t_stage_host.p;
  char* dbase = (char*)ctx->t_arena.p;
  double* T0 = (double*)(hbase + ctx->t_geo_bytes);
  float* a0 = (float*)(T0 + 12 * B);
  const size_t g0 = ctx->t_goff[b0], g1 = (b1 < B) ? ctx->t_goff[b1] : ctx->t_geo_bytes;
  const TrackJobDev* hj = (const TrackJobDev*)ctx->t_jobs_host.p;
  if (!ctx->t_direct) CU(cudaMemcpyAsync(dbase + g0, hbase + g0, g1 - g0, cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync((double*)ctx->t_T0.p + 12 * b0, T0 + 12 * b0, sizeof(double) * 12 * n, cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync((float*)ctx->t_a0.p + b0, a0 + b0, sizeof(float) * n, cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync((TrackJobDev*)ctx->t_jobs_dev.p + b0, hj + b0, sizeof(TrackJobDev) * n, cudaMemcpyHostToDevice, stream));
  return HSO_OK;
}

// Host workers that flatten the jobs of a batch in order; wait_chunk(c) returns once every job of chunk c is staged. Flattening the
// reference's Feature lists to SoA is host work of the boundary; it runs ahead of the copies and launches the caller enqueues.
struct StageWorkers {
  hso_ctx* ctx; const hso_track_job* jobs; int B;
  std::vector<int> bounds;    // chunk c = jobs [bounds[c], bounds[c + 1])
  std::vector<int> chunk_of;  // job -> chunk
  std::vector<std::atomic<int>> done;
  bool inline_only;
  int inline_next = 0;
  StageWorkers(hso_ctx* c, const hso_track_job* j, int B_, const std::vector<int>& bounds_)
      : ctx(c), jobs(j), B(B_), bounds(bounds_), chunk_of(B_), done(bounds_.size() - 1) {
    for (auto& d : done) d.store(0);
    for (size_t k = 0; k + 1 < bounds.size(); ++k)
      for (int b = bounds[k]; b < bounds[k + 1]; ++b) chunk_of[b] = (int)k;
    inline_only = B < 16;  // a small batch is flattened by the caller's thread: waking the pool costs more than the work
    if (!inline_only) {
      if (ctx->pool.size() == 0) ctx->pool.start(flatten_threads_default());
      ctx->pool.submit(&StageWorkers::item, this, B);
    }
  }
  static void item(void* self, int b) {
    StageWorkers* w = (StageWorkers*)self;
    track_stage_one(w->ctx, w->jobs, b);
    w->done[w->chunk_of[b]].fetch_add(1, std::memory_order_release);
  }
  void wait_chunk(int c) {
    const int want = bounds[c + 1] - bounds[c];
    if (inline_only) {
      for (; inline_next < bounds[c + 1]; ++inline_next) item(this, inline_next);
      return;
    }
    while (done[c].load(std::memory_order_acquire) < want) std::this_thread::yield();
  }
  ~StageWorkers() { if (!inline_only) ctx->pool.wait(); }  // every job is staged (or being staged by the caller) before the records go away
};

int hso_track_stage(hso_ctx* ctx, const hso_track_params* prm, int B, const hso_track_job* jobs, int trace_cap) {
  int rc = track_plan(ctx, prm, B, jobs, trace_cap);
  if (rc != HSO_OK) return rc;
  if (ctx->t_direct) {
    rc = track_copy_raw(ctx, jobs, 0, B, ctx->stream);
    if (rc == HSO_OK) rc = track_copy_range(ctx, 0, B, ctx->stream);
    if (rc != HSO_OK) return rc;
    CU(launch_track_compact((TrackJobDev*)ctx->t_jobs_dev.p, B, ctx->stream, &ctx->launches));
    CU(cudaStreamSynchronize(ctx->stream));  // the caller's arrays may change once this call returns
    return HSO_OK;
  }
  {
    StageWorkers w(ctx, jobs, B, std::vector<int>{0, B});
    w.wait_chunk(0);
  }
  return track_copy_range(ctx, 0, B, ctx->stream);
}

int hso_track_restage_frames(hso_ctx* ctx, int B, const hso_frame_id* ref, const hso_frame_id* cur) {
  if (!ctx || B != ctx->tB || !ref || !cur) return HSO_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  TrackJobDev* hj = (TrackJobDev*)ctx->t_jobs_host.p;  // the plan's host mirror; the copy below reads a snapshot of it
  for (int b = 0; b < B; ++b) {
    if (!get_frame(ctx, ref[b]) || !get_frame(ctx, cur[b])) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown frame id");
    hj[b].ref_pyr = get_frame(ctx, ref[b])->pyr;
    hj[b].cur_pyr = get_frame(ctx, cur[b])->pyr;
  }
  // only the two pyramid pointers change: patch those 16 bytes of every device record (the rest — e.g. F, which the device wrote itself in
  // direct-input mode — stays as it is)
  static_assert(offsetof(TrackJobDev, ref_pyr) == 0 && offsetof(TrackJobDev, cur_pyr) == sizeof(void*), "pyramid pointers lead the record");
  void* region = nullptr;
  CU(ctx->t_jobs_ring.acquire(2 * sizeof(void*) * B, &region));
  const uint8_t** pp = (const uint8_t**)region;
  for (int b = 0; b < B; ++b) { pp[2 * b] = hj[b].ref_pyr; pp[2 * b + 1] = hj[b].cur_pyr; }
  CU(cudaMemcpy2DAsync(ctx->t_jobs_dev.p, sizeof(TrackJobDev), region, 2 * sizeof(void*), 2 * sizeof(void*), B, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx->t_jobs_ring.commit(ctx->stream));
  return HSO_OK;
}

// Per-level kernel timing (CUDA events on the launching stream) for bench.py's roofline: events bracket each k_track_level launch.
static int track_profile_flush(hso_ctx* ctx) {
  if (!ctx->t_prof_pending) return HSO_OK;
  const hso_track_params& prm = ctx->tprm;
  const int n = prm.max_level - prm.min_level + 1;
  CU(cudaEventSynchronize(ctx->t_ev[n]));
  int slot = 0;
  for (int level = prm.max_level; level >= prm.min_level; --level, ++slot) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->t_ev[slot], ctx->t_ev[slot + 1]));
    ctx->t_level_ms[level] += ms;
    ctx->t_level_launches[level] += 1;
  }
  ctx->t_prof_pending = 0;
  return HSO_OK;
}

int hso_track_set_profile(hso_ctx* ctx, int on) {
  if (!ctx) return HSO_ERR_INVALID;
  if (on) {
    for (int i = 0; i <= kMaxLevels; ++i)
      if (!ctx->t_ev[i]) CU(cudaEventCreate(&ctx->t_ev[i]));
    for (int l = 0; l < kMaxLevels; ++l) { ctx->t_level_ms[l] = 0; ctx->t_level_launches[l] = 0; }
    ctx->t_prof_pending = 0;
  }
  ctx->t_profile = on ? 1 : 0;
  return HSO_OK;
}

int hso_track_level_profile(hso_ctx* ctx, int level, double* ms_total, uint64_t* launches) {
  if (!ctx || level < 0 || level >= kMaxLevels) return HSO_ERR_INVALID;
  int rc = track_profile_flush(ctx);
  if (rc != HSO_OK) return rc;
  if (ms_total) *ms_total = ctx->t_level_ms[level];
  if (launches) *launches = ctx->t_level_launches[level];
  return HSO_OK;
}

// init + one launch per level + finish for jobs [b0, b0 + B) of the staged batch (B = problems in these launches: decides the launch shape)
// shape_B: problems that share the GPU at the same time (the whole batch of a pipelined call), stream: where to launch
static int track_run_range(hso_ctx* ctx, int b0, int B, bool profile, int shape_B = 0, cudaStream_t stream = nullptr) {
  const hso_track_params& prm = ctx->tprm;
  if (!stream) stream = ctx->stream;
  if (shape_B <= 0) shape_B = B;
  const TrackJobDev* jd = (const TrackJobDev*)ctx->t_jobs_dev.p + b0;
  if (profile) {
    int rc = track_profile_flush(ctx);
    if (rc != HSO_OK) return rc;
  }
  CU(launch_track_init(jd, (const double*)ctx->t_T0.p + 12 * b0, (const float*)ctx->t_a0.p + b0, B, stream, &ctx->launches));
  int slot = 0;
  for (int level = prm.max_level; level >= prm.min_level; --level) {
    if (profile) CU(cudaEventRecord(ctx->t_ev[slot++], stream));
    TrackLevelParams p;
    memset(&p, 0, sizeof p);
    p.ic = prm.inverse_comp; p.max_level = prm.max_level; p.level = level; p.n_iter = prm.n_iter;
    p.trace_cap = ctx->t_trace_cap;
    p.w = ctx->geom.w[level]; p.h = ctx->geom.h[level];
    p.level_off = ctx->geom.off[level];
    p.img_bytes = ctx->geom.stage_bytes[level];
    p.cam = ctx->camdev;
    // Launch shape of this level. FAST keeps the level image and the CTA's reference-patch cache in shared memory: take the
    // smallest cluster size whose per-CTA share fits 227 KB, but never fewer CTAs than it takes to cover the 148 SMs when the
    // batch is small (single-problem latency mode). Otherwise fall back to the global-memory path.
    const int maxF = std::max(ctx->t_maxF, 1);
    const int f_cluster = ctx->t_shape[level][0] ? ctx->t_shape[level][0] : ctx->t_cluster;
    const int f_threads = ctx->t_shape[level][1] ? ctx->t_shape[level][1] : ctx->t_threads;
    int c_min = f_cluster;
    if (c_min == 0) {
      c_min = shape_B >= 148 ? 1 : (shape_B >= 74 ? 2 : (shape_B >= 37 ? 4 : 8));
      // ... but never more CTAs than it takes to give every thread one patch (a cluster only adds barriers beyond that)
      int c_work = 1;
      while (c_work < 8 && c_work * 512 < maxF) c_work *= 2;
      c_min = std::min(c_min, c_work);
    }
    int cluster = 0, threads = 0;
    bool pair_ctas = false;
    // Two or more problems per SM in THIS launch (a chunk of the pipelined call has less than one per SM, and a 256-thread CTA alone on its SM
    // is just slower): two CTAs of 256 threads share an SM when their footprint is under half an SM (the |r| scratch then lives in global
    // memory) — one problem's serial control step, barriers and staging latencies overlap the other's evaluation.
    const bool pairs_ok = !ctx->t_no_pair && !f_threads && c_min == 1 && B >= 2 * 148 && maxF > 256;
    if (pairs_ok && !ctx->t_no_stream) {
      // streamed-cache mode (image + 8-warp ring + histogram). Forward, measured at B = 1184, F = 3000 against one 512-thread CTA with the cache
      // resident: level 4 -7 %, level 3 -3 %, level 2 -3 % (profiles/r2w_shapes.txt); level 1 (77 KB image) does not fit twice.
      TrackLevelParams q = p;
      q.fast = 3; q.cluster = 1; q.hist_bits = 11; q.absres_smem = 0;
      q.pc = (maxF + 255) / 256 * 256;
      // two CTAs fit an SM when each needs at most (228 KB - 2 x 1 KB reserved) / 2 = 113 KB
      const size_t pair_limit = (228 * 1024 - 2 * 1024) / 2;
      if (track_level_smem_bytes(q, 256) <= pair_limit) { p.fast = 3; p.pc = q.pc; p.cluster = 1; p.hist_bits = 11; cluster = 1; threads = 256; pair_ctas = true; }
      // (level 1 at 640x480 fits twice only with the single-buffered ring — 77 KB image + 21 KB ring + 14 KB; measured, no gain over one 512-thread
      // CTA with the double-buffered ring: 3.94 vs 3.98 ms at B = 2368)
    }
    if (!cluster && prm.inverse_comp && !ctx->t_no_stream) {
      // inverse-compositional, cached reference intensities + gradients (what the reference precomputes, src/CoarseTracker.cpp:482-492) streamed
      // from L2: ~48 instructions per residual term against ~73 for the dual-image mode below, which re-derives them from the reference image
      const int cc = c_min;
      const int th = f_threads ? f_threads : std::min(512, std::max(64, ((maxF + cc - 1) / cc + 31) / 32 * 32));
      const int kpt = (maxF + cc * th - 1) / (cc * th);
      p.fast = 3; p.pc = kpt * th; p.cluster = cc; p.hist_bits = 11;
      if (track_level_smem_bytes(p, th) <= 227 * 1024) { cluster = cc; threads = th; }
      else if (!ctx->t_no_ring1) {
        // level 1 at 640x480: 16 warps x 2 buffers x 63 rows = 258 KB do not fit, one buffer per warp (129 KB) beside the 77 KB image does (mode 4):
        // 2.28 ms against 2.96 ms for the dual-image mode at B = 1184, F = 3000
        p.fast = 4;
        if (track_level_smem_bytes(p, th) <= 227 * 1024) { cluster = cc; threads = th; }
      }
    }
    if (!cluster && prm.inverse_comp && !ctx->t_no_dual) {
      // inverse-compositional: keep BOTH levels resident and recompute the reference samples per evaluation — no F-dependent
      // cache, so one CTA per problem fits whenever two copies of the level do
      const int cc = c_min;
      int th = f_threads ? f_threads : std::min(512, std::max(64, ((maxF + cc - 1) / cc + 31) / 32 * 32));
      // pairs of 256-thread CTAs, measured at B = 1184, F = 3000: level 4 -12 %, level 3 -5 %, level 2 -4 % (profiles/r2m_ic_shapes.txt)
      bool pair = false;
      if (pairs_ok && th > 256) {
        TrackLevelParams q = p;
        q.fast = 2; q.cluster = 1; q.hist_bits = 11; q.absres_smem = 0;
        q.pc = (maxF + 255) / 256 * 256;
        if (track_level_smem_bytes(q, 256) <= (228 * 1024 - 2 * 1024) / 2) { th = 256; pair = true; }
      }
      const int kpt = (maxF + cc * th - 1) / (cc * th);
      p.fast = 2; p.pc = kpt * th; p.cluster = cc; p.hist_bits = 11;
      if (track_level_smem_bytes(p, th) <= 227 * 1024) { cluster = cc; threads = th; pair_ctas = pair; }
    }
    for (int cc = c_min; cc <= 8 && !cluster; cc *= 2) {
      int th = f_threads ? f_threads : std::min(512, std::max(64, ((maxF + cc - 1) / cc + 31) / 32 * 32));
      const int kpt = (maxF + cc * th - 1) / (cc * th);
      p.pc = kpt * th; p.cluster = cc;
      if (ctx->t_force_stream) {  // tuning / tests: the streamed-cache mode wherever it fits
        p.fast = 3; p.hist_bits = 11;
        if (track_level_smem_bytes(p, th) <= 227 * 1024) { cluster = cc; threads = th; break; }
      }
      p.fast = 1;
      p.hist_bits = 11;
      if (track_level_smem_bytes(p, th) <= 227 * 1024) { cluster = cc; threads = th; break; }
      p.hist_bits = 8;
      if (track_level_smem_bytes(p, th) <= 227 * 1024) { cluster = cc; threads = th; break; }
      // image resident, reference-patch cache streamed from L2 through a per-warp ring (mode 3) before the problem is split over more
      // CTAs — a cluster costs two cluster barriers per trial, a second staged image and a second (redundant) control step
      if (!ctx->t_no_stream) {
        p.fast = 3; p.hist_bits = 11;
        if (track_level_smem_bytes(p, th) <= 227 * 1024) { cluster = cc; threads = th; break; }
      }
      if (f_cluster) break;  // the caller fixed the cluster size
    }
    if (!cluster) {
      cluster = c_min;
      threads = f_threads ? f_threads : std::min(512, std::max(64, ((maxF + cluster - 1) / cluster + 31) / 32 * 32));
      p.fast = 0; p.pc = 0; p.hist_bits = 11; p.cluster = cluster;
    }
    if (p.fast) {  // keep the |r| scratch of the threshold selection in shared memory too when it still fits
      p.absres_smem = 1;
      if (track_level_smem_bytes(p, threads) > 227 * 1024 || ctx->t_abs_global || pair_ctas) p.absres_smem = 0;
    }
    ctx->t_used[level][0] = cluster; ctx->t_used[level][1] = threads; ctx->t_used[level][2] = p.fast; ctx->t_used[level][3] = p.absres_smem;
    ctx->t_used[level][4] = p.hist_bits;
    CU(launch_track_level(p, jd, B, cluster, threads, stream, &ctx->launches));
  }
  if (profile) {
    CU(cudaEventRecord(ctx->t_ev[slot], stream));
    ctx->t_prof_pending = 1;
  }
  CU(launch_track_finish(jd, (hso_track_result*)ctx->t_out_dev.p + b0, B, stream, &ctx->launches));
  return HSO_OK;
}

int hso_track_run(hso_ctx* ctx) {
  if (!ctx || ctx->tB <= 0) return HSO_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  return track_run_range(ctx, 0, ctx->tB, ctx->t_profile != 0);
}

int hso_track_collect(hso_ctx* ctx, hso_track_result* out, hso_trace* trace, int* trace_len) {
  if (!ctx || ctx->tB <= 0 || !out) return HSO_ERR_INVALID;
  const int B = ctx->tB;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(ctx->t_out_host.p, ctx->t_out_dev.p, sizeof(hso_track_result) * B, cudaMemcpyDeviceToHost, ctx->stream));
  if (trace && ctx->t_trace_cap) {
    for (int b = 0; b < B; ++b)
      CU(cudaMemcpyAsync(trace + (size_t)b * ctx->t_trace_cap, (char*)ctx->t_arena.p + ctx->t_trace_off[b], sizeof(hso_trace) * ctx->t_trace_cap,
                         cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->t_out_host.p, sizeof(hso_track_result) * B);
  if (trace_len) for (int b = 0; b < B; ++b) trace_len[b] = out[b].trace_len;
  return HSO_OK;
}

int hso_coarse_track_batch(hso_ctx* ctx, const hso_track_params* prm, int B, const hso_track_job* jobs, hso_track_result* out, hso_trace* trace,
                           int trace_cap, int* trace_len) {
  if (!ctx) return HSO_ERR_INVALID;
  int rc = hso_track_stage(ctx, prm, B, jobs, trace ? trace_cap : 0);
  if (rc != HSO_OK) return rc;
  StageTimer tm(ctx, 1);
  rc = hso_track_run(ctx);
  if (rc != HSO_OK) return rc;
  rc = hso_track_collect(ctx, out, trace, trace_len);
  if (rc != HSO_OK) return rc;
  tm.stop_after_sync();
  // CoarseTracker::run returns 0 and leaves the pose untouched when the reference frame has no features (:53)
  for (int b = 0; b < B; ++b)
    if (jobs[b].n_features == 0) {
      memcpy(out[b].T_cur_ref, jobs[b].T_cur_ref, sizeof(double) * 12);
      if (jobs[b].exposure_rat >= 0) out[b].exposure_rat = jobs[b].exposure_rat;  // else: the ratio of the two frames' statistics, formed on the device
      out[b].n_tracked = 0;
    }
  return HSO_OK;
}

int hso_coarse_track(hso_ctx* ctx, const hso_track_params* prm, const hso_track_job* job, hso_track_result* out, hso_trace* trace, int trace_cap,
                     int* trace_len) {
  return hso_coarse_track_batch(ctx, prm, 1, job, out, trace, trace_cap, trace_len);
}

// ---- F1 + F2, chunk-pipelined ------------------------------------------------------------------------------------------------
// The front end of FrameHandlerMono::addImage for B independent streams: new Frame(cam, img) (src/frame_handler_mono.cpp:92) then
// CoarseTracker::run(ref, new_frame) (:190-204). The batch is cut into chunks; while chunk c's pyramid and tracker kernels run on
// the context stream, chunk c+1's images and feature arrays are flattened on the host and copied on a second stream.
int hso_add_frames_track_batch(hso_ctx* ctx, const hso_track_params* prm, int B, const uint8_t* const* imgs, int W, int H, int stride,
                               const hso_track_job* jobs_in, hso_frame_id* new_ids, float* integral, float* grad_mean, hso_track_result* out) {
  if (!ctx || !prm || B <= 0 || !imgs || !jobs_in || !new_ids || !out) return HSO_ERR_INVALID;
  if (W != ctx->cam.width || H != ctx->cam.height || stride < W) return fail(ctx, HSO_ERR_INVALID, "image size does not match the camera model");
  CU(cudaSetDevice(ctx->device));
  // tuning aid: HSO_PIPE_TRACE=1 prints host-side time stamps of the call's phases (ms since entry) to stderr
  const bool ptrace = getenv("HSO_PIPE_TRACE") != nullptr;
  const auto t_entry = std::chrono::steady_clock::now();
  auto stamp = [&](const char* what) {
    if (ptrace) fprintf(stderr, "[pipe] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count());
  };
  for (int i = 0; i < B; ++i)
    if (!imgs[i]) return fail(ctx, HSO_ERR_INVALID, "null image");
  for (int i = 0; i < B; ++i) {
    int rc = alloc_frame(ctx, &new_ids[i]);
    if (rc != HSO_OK) { for (int j = 0; j < i; ++j) unuse_frame(ctx, new_ids[j]); return rc; }
  }
  auto release_all = [&]() { for (int i = 0; i < B; ++i) unuse_frame(ctx, new_ids[i]); };
  std::vector<cudaEvent_t> tr_ev;
  std::vector<hso_track_job> jobs(jobs_in, jobs_in + B);
  for (int b = 0; b < B; ++b) jobs[b].cur = new_ids[b];
  stamp("frames allocated");
  int rc = track_plan(ctx, prm, B, jobs.data(), 0);  // synchronises ctx->stream: nothing of a previous call is in flight below
  if (rc != HSO_OK) { release_all(); return rc; }
  stamp("track_plan");
  // Everything that can fail after the frames were allocated runs inside `pipeline`, so that every error path — a CUDA error in the middle of
  // the chunk loop included — goes through the same clean-up: wait for the streams, hand the new frames back.
  int S = 1;
  StageTimer tm(ctx, 1);
  std::unique_ptr<StageWorkers> workers;  // started right after the first chunk's image copy is on its way (spawning the pool takes ~0.3 ms)
  auto pipeline = [&]() -> int {
  if (!ctx->copy_stream) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  // chunk = one wave of single-CTA problems (148 SMs); few enough chunks that the per-chunk launch overhead stays small
  // one wave of single-CTA problems on four streams: measured best with this round's kernels and the compact feature layout (e2e at B = 2368:
  // 111:3 3.28, 148:3 3.27, 111:4 3.35, 74:4 3.30, 185:3 3.32, 148:4 3.36 M it/s)
  const int unit = ctx->pipe_chunk > 0 ? ctx->pipe_chunk : 148;
  int chunk = B <= unit ? B : unit;
  while ((B + chunk - 1) / chunk > 32) chunk += unit;
  // the first two chunks ramp up (1/3, 2/3 of a chunk): the GPU starts after a third of a chunk's flattening + copy instead of a whole one;
  // the last two ramp down the same way: what is left to compute after the final copy is a third of a chunk (the call is copy bound)
  std::vector<int> bounds{0};
  const bool ramp = B > 4 * chunk && chunk >= 12;
  if (ramp) { bounds.push_back(chunk / 3); bounds.push_back(chunk / 3 + 2 * chunk / 3); }
  const int tail_sz = ramp ? (chunk / 3 + 2 * chunk / 3) : 0;
  while (bounds.back() < B - tail_sz) bounds.push_back(std::min(B - tail_sz, bounds.back() + chunk));
  if (ramp) { bounds.push_back(B - chunk / 3); bounds.push_back(B); }
  const int n_chunks = (int)bounds.size() - 1;
  S = std::max(1, std::min(n_chunks, ctx->pipe_n_streams > 0 ? ctx->pipe_n_streams : 4));
  while ((int)ctx->pipe_streams.size() < S - 1) {
    cudaStream_t st; cudaEvent_t e;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->pipe_streams.push_back(st);
    ctx->pipe_ev.push_back(e);
  }
  while ((int)ctx->chunk_ev.size() < n_chunks) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->chunk_ev.push_back(e);
  }
  // every record array is reserved for the whole batch up front: chunks must not re-allocate what an earlier chunk still uses
  CU(ctx->pyr_jobs_host.reserve(sizeof(PyrJobDev) * B));
  CU(ctx->pyr_jobs_dev.reserve(sizeof(PyrJobDev) * B));
  if (ctx->pyr_counters.cap < sizeof(unsigned) * (size_t)B) {
    CU(ctx->pyr_counters.reserve(sizeof(unsigned) * B));
    CU(cudaMemsetAsync(ctx->pyr_counters.p, 0, ctx->pyr_counters.cap, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  // tuning aid (tools/e2e_breakdown.py): HSO_PIPE_DEBUG bit 0 skips the kernel launches, bit 1 the image / feature copies (bit 2: images only, bit 3: features only), to time the stages alone
  const char* dbg_env = getenv("HSO_PIPE_DEBUG");
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  std::vector<const uint8_t*> srcs(B);
  // direct-input mode: every small record of the batch (pyramid jobs, tracker jobs, initial poses, exposure ratios: ~300 B per problem) goes out
  // in four copies up front instead of five per chunk — each copy costs the copy engine a few microseconds whatever its size, and the call is
  // bound by the copy stream
  const bool records_up_front = ctx->t_direct && !(dbg & (2 | 8));
  if (records_up_front) {
    for (int i = 0; i < B; ++i) srcs[i] = get_frame(ctx, new_ids[i])->pyr + ctx->geom.off[0];
    rc = run_pyramid(ctx, B, new_ids, srcs.data(), W, 1, 0, ctx->copy_stream);
    if (rc != HSO_OK) return rc;
    track_fill_pose(ctx, jobs.data(), 0, B);
    rc = track_copy_range(ctx, 0, B, ctx->copy_stream);
    if (rc != HSO_OK) return rc;
  }
  if (ptrace) {  // timed events: copy end, first kernel start, last kernel end of every chunk
    tr_ev.resize(3 * n_chunks + 1);
    for (auto& e : tr_ev) cudaEventCreate(&e);
    cudaEventRecord(tr_ev[3 * n_chunks], ctx->copy_stream);
  }
  for (int c = 0; c < n_chunks; ++c) {
    const int b0 = bounds[c], b1 = bounds[c + 1], n = b1 - b0;
    // images straight into the level-0 slots (one 2-D copy for equally spaced images going to consecutive slots)
    bool one_copy = n > 1 && stride == W;
    const ptrdiff_t spacing = n > 1 ? imgs[b0 + 1] - imgs[b0] : 0;
    for (int i = b0 + 1; i < b1 && one_copy; ++i) one_copy = (imgs[i] - imgs[i - 1] == spacing) && (new_ids[i] == new_ids[i - 1] + 1);
    one_copy = one_copy && spacing >= (ptrdiff_t)W * H;
    if (one_copy && !(dbg & 2) && !(dbg & 4))
      CU(cudaMemcpy2DAsync(get_frame(ctx, new_ids[b0])->pyr + ctx->geom.off[0], ctx->pyr_slot_bytes, imgs[b0], (size_t)spacing, (size_t)W * H, n,
                           cudaMemcpyHostToDevice, ctx->copy_stream));
    for (int i = b0; i < b1; ++i) {
      FrameSlot* s = get_frame(ctx, new_ids[i]);
      if (!one_copy && !(dbg & 2) && !(dbg & 4)) CU(cudaMemcpy2DAsync(s->pyr + ctx->geom.off[0], W, imgs[i], stride, W, H, cudaMemcpyHostToDevice, ctx->copy_stream));
      srcs[i] = s->pyr + ctx->geom.off[0];
    }
    if (!ctx->t_direct && !workers) workers.reset(new StageWorkers(ctx, jobs.data(), B, bounds));
    if (!records_up_front) {
      rc = run_pyramid(ctx, n, new_ids + b0, srcs.data() + b0, W, 1, b0, ctx->copy_stream);
      if (rc != HSO_OK) return rc;
    }
    if (ctx->t_direct) {
      rc = (dbg & (2 | 8)) ? HSO_OK : track_copy_raw(ctx, jobs.data(), b0, b1, ctx->copy_stream, !records_up_front);
      if (rc != HSO_OK) return rc;
    } else {
      workers->wait_chunk(c);
    }
    rc = ((dbg & (2 | 8)) || records_up_front) ? HSO_OK : track_copy_range(ctx, b0, b1, ctx->copy_stream);
    if (rc != HSO_OK) return rc;
    CU(cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream));
    if (ptrace) cudaEventRecord(tr_ev[3 * c], ctx->copy_stream);
    cudaStream_t cs = (c % S == 0) ? ctx->stream : ctx->pipe_streams[c % S - 1];
    CU(cudaStreamWaitEvent(cs, ctx->chunk_ev[c], 0));
    if (dbg & 1) continue;
    if (ptrace) cudaEventRecord(tr_ev[3 * c + 1], cs);
    if (ctx->t_direct) CU(launch_track_compact((TrackJobDev*)ctx->t_jobs_dev.p + b0, n, cs, &ctx->launches));
    CU(launch_pyramid(ctx->geom, (const PyrJobDev*)ctx->pyr_jobs_dev.p + b0, n, W, ctx->resize_tabs.data(), ctx->cfg.materialize_sobel,
                      (unsigned*)ctx->pyr_counters.p + b0, 1, cs, &ctx->launches));
    // (measured, profiles/r2t_tail_sweep.txt: giving the last chunks latency shapes — clusters of 2/4/8 CTAs by the problems still to come — and a
    // geometric ramp-down on streams of their own moves the end of the call by 0.15 ms of 12: the SMs are held by the older chunks' CTAs,
    // and a problem's four level launches take >= 1.3 ms whatever its chunk size)
    rc = track_run_range(ctx, b0, n, false, B, cs);
    if (rc != HSO_OK) return rc;
    if (ptrace) cudaEventRecord(tr_ev[3 * c + 2], cs);
    if (ptrace && (c < 3 || c + 2 >= n_chunks)) { char nm[48]; snprintf(nm, sizeof nm, "chunk %d (%d jobs) enqueued", c, n); stamp(nm); }
  }
    if (ptrace) { cudaStreamSynchronize(ctx->copy_stream); stamp("copy stream drained"); }
    return HSO_OK;
  };
  rc = pipeline();
  stamp("pipeline enqueued");
  workers.reset();  // joins the flattening threads
  for (int k = 0; k < S - 1 && k < (int)ctx->pipe_streams.size(); ++k) {  // join the extra compute streams into the context stream
    cudaEventRecord(ctx->pipe_ev[k], ctx->pipe_streams[k]);
    cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[k], 0);
  }
  if (rc != HSO_OK) {
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    release_all();
    return rc;
  }
  if (ptrace) {
    cudaStreamSynchronize(ctx->stream);
    stamp("kernels drained");
    const int nc = ((int)tr_ev.size() - 1) / 3;
    for (int c = 0; c < nc; ++c) {
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, tr_ev[3 * nc], tr_ev[3 * c]); cudaEventElapsedTime(&b, tr_ev[3 * nc], tr_ev[3 * c + 1]); cudaEventElapsedTime(&d, tr_ev[3 * nc], tr_ev[3 * c + 2]);
      fprintf(stderr, "[pipe] chunk %2d: copy done %7.3f  kernels %7.3f .. %7.3f ms\n", c, a, b, d);
    }
    for (auto& e : tr_ev) cudaEventDestroy(e);
  }
  rc = hso_track_collect(ctx, out, nullptr, nullptr);
  if (rc != HSO_OK) return rc;
  stamp("results collected");
  rc = read_stats(ctx, B, new_ids, integral, grad_mean);
  if (rc != HSO_OK) return rc;
  stamp("statistics read");
  tm.stop_after_sync();
  for (int b = 0; b < B; ++b)
    if (jobs[b].n_features == 0) {  // CoarseTracker::run returns 0 and leaves the pose untouched when the reference has no features (:53)
      memcpy(out[b].T_cur_ref, jobs[b].T_cur_ref, sizeof(double) * 12);
      if (jobs[b].exposure_rat >= 0) out[b].exposure_rat = jobs[b].exposure_rat;
      out[b].n_tracked = 0;
    }
  return HSO_OK;
}

// ---- F3-inner ---------------------------------------------------------------------------------------------------------------
int hso_align_batch(hso_ctx* ctx, hso_frame_id cur, int M, const hso_align_job* jobs, const hso_frame_id* ref_frames, int align_max_iter,
                    hso_align_result* out) {
  if (!ctx || M < 0 || (M > 0 && (!jobs || !ref_frames || !out))) return HSO_ERR_INVALID;
  if (M == 0) return HSO_OK;
  FrameSlot* fc = get_frame(ctx, cur);
  if (!fc) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown current frame id");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));  // the pinned staging below is reused between calls
  CU(ctx->a_jobs_host.reserve(sizeof(AlignJobDev) * M));
  CU(ctx->a_jobs_dev.reserve(sizeof(AlignJobDev) * M));
  CU(ctx->a_out_dev.reserve(sizeof(hso_align_result) * M));
  CU(ctx->a_out_host.reserve(sizeof(hso_align_result) * M));
  AlignJobDev* hj = (AlignJobDev*)ctx->a_jobs_host.p;
  const int max_search = std::min(ctx->cfg.n_pyr_levels, ctx->geom.n_levels) - 1;  // getBestSearchLevel(A, nPyrLevels-1)
  for (int m = 0; m < M; ++m) {
    const hso_align_job& j = jobs[m];
    FrameSlot* fr = get_frame(ctx, ref_frames[m]);
    if (!fr) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown reference frame id in align job");
    if (j.ref_level < 0 || j.ref_level >= ctx->geom.n_levels || j.search_level < 0 || j.search_level > max_search)
      return fail(ctx, HSO_ERR_INVALID, "align job level out of range");
    if (j.type == 1 && !fc->sobel) return fail(ctx, HSO_ERR_INVALID, "edgelet jobs need hso_cfg.materialize_sobel (checkNormal reads sobelX_/Y_)");
    hj[m].job = j;
    hj[m].ref_pyr = fr->pyr;
  }
  StageTimer tm(ctx, 2);
  CU(cudaMemcpyAsync(ctx->a_jobs_dev.p, hj, sizeof(AlignJobDev) * M, cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_align(ctx->geom, fc->pyr, fc->sobel, (const AlignJobDev*)ctx->a_jobs_dev.p, M, align_max_iter, (hso_align_result*)ctx->a_out_dev.p,
                  ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(ctx->a_out_host.p, ctx->a_out_dev.p, sizeof(hso_align_result) * M, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->a_out_host.p, sizeof(hso_align_result) * M);
  tm.stop_after_sync();
  return HSO_OK;
}

// ---- N1 ---------------------------------------------------------------------------------------------------------------------
int hso_reproject_match(hso_ctx* ctx, hso_frame_id cur, const double T_cur_w[12], int n_poses, const double* T_f_w, int M,
                        const hso_reproj_cand* cands, const hso_reproj_grid* grid, const int32_t* cell_order, hso_reproj_result* out,
                        hso_reproj_summary* summary) {
  if (!ctx || !T_cur_w || !grid || !summary || M < 0 || n_poses < 0) return HSO_ERR_INVALID;
  memset(summary, 0, sizeof *summary);
  const int n_cells = grid->n_cols * grid->n_rows;
  if (grid->cell_size <= 0 || grid->n_cols <= 0 || grid->n_rows <= 0 || n_cells > 4096 || grid->max_fts < 0 || !cell_order)
    return fail(ctx, HSO_ERR_INVALID, "bad reprojection grid (cells must be <= 4096)");
  if ((long long)grid->n_cols * grid->cell_size < ctx->cam.width || (long long)grid->n_rows * grid->cell_size < ctx->cam.height)
    return fail(ctx, HSO_ERR_INVALID, "reprojection grid does not cover the image");
  if (M == 0) { summary->used_cell_all = 1; return HSO_OK; }
  if (!cands || !out || !T_f_w || n_poses == 0) return HSO_ERR_INVALID;
  if (M > 16384) return fail(ctx, HSO_ERR_CAPACITY, "more than 16384 reprojection candidates");
  FrameSlot* fc = get_frame(ctx, cur);
  if (!fc) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown current frame id");
  {
    std::vector<uint8_t> seen(n_cells, 0);
    for (int i = 0; i < n_cells; ++i) {
      if (cell_order[i] < 0 || cell_order[i] >= n_cells || seen[cell_order[i]]) return fail(ctx, HSO_ERR_INVALID, "cell_order is not a permutation");
      seen[cell_order[i]] = 1;
    }
  }
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  // staging blob: [cands | ref_pyr pointers | poses | cell_order], then device-only [align jobs | align results | results | summary]
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = (o + 127) / 128 * 128; o = r + bytes; return r; };
  const size_t o_c = take(sizeof(hso_reproj_cand) * M), o_rp = take(sizeof(void*) * M), o_T = take(sizeof(double) * 12 * n_poses),
               o_co = take(sizeof(int32_t) * n_cells);
  const size_t staged = o;
  const size_t o_j = take(sizeof(AlignJobDev) * M), o_ar = take(sizeof(hso_align_result) * M), o_res = take(sizeof(hso_reproj_result) * M),
               o_sum = take(sizeof(hso_reproj_summary));
  CU(ctx->r_arena.reserve(o));
  CU(ctx->r_stage_host.reserve(staged));
  CU(ctx->r_out_host.reserve(sizeof(hso_reproj_result) * M + sizeof(hso_reproj_summary)));
  char* h = (char*)ctx->r_stage_host.p;
  char* d = (char*)ctx->r_arena.p;
  memcpy(h + o_c, cands, sizeof(hso_reproj_cand) * M);
  const uint8_t** rp = (const uint8_t**)(h + o_rp);
  const int max_search = std::min(ctx->cfg.n_pyr_levels, ctx->geom.n_levels) - 1;  // getBestSearchLevel(A, nPyrLevels-1)
  for (int i = 0; i < M; ++i) {
    const hso_reproj_cand& c = cands[i];
    rp[i] = nullptr;
    if (c.host_pose < 0 || c.host_pose >= n_poses || c.ref_pose >= n_poses) return fail(ctx, HSO_ERR_INVALID, "pose index out of range in reprojection candidate");
    if (c.pt_type < 0 || c.pt_type > 4 || c.pt_ftr_type < 0 || c.pt_ftr_type > 2) return fail(ctx, HSO_ERR_INVALID, "bad point type in reprojection candidate");
    if (c.ref_pose >= 0) {
      FrameSlot* fr = get_frame(ctx, c.ref_frame);
      if (!fr) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown reference frame id in reprojection candidate");
      if (c.ref_level < 0 || c.ref_level >= ctx->geom.n_levels) return fail(ctx, HSO_ERR_INVALID, "reference level out of range");
      if (c.ftr_type == 1 && !fc->sobel) return fail(ctx, HSO_ERR_INVALID, "edgelet candidates need hso_cfg.materialize_sobel (checkNormal reads sobelX_/Y_)");
      rp[i] = fr->pyr;
    }
  }
  memcpy(h + o_T, T_f_w, sizeof(double) * 12 * n_poses);
  memcpy(h + o_co, cell_order, sizeof(int32_t) * n_cells);
  StageTimer tm(ctx, ST_REPROJECT);
  CU(cudaMemcpyAsync(d, h, staged, cudaMemcpyHostToDevice, ctx->stream));
  ReprojKParams kp;
  memset(&kp, 0, sizeof kp);
  kp.cam = ctx->camdev;
  memcpy(kp.T_cur_w, T_cur_w, sizeof kp.T_cur_w);
  kp.T_f_w = (const double*)(d + o_T);
  kp.n_poses = n_poses; kp.M = M; kp.cell_size = grid->cell_size; kp.n_cols = grid->n_cols; kp.max_search_level = max_search;
  CU(launch_reproject(kp, (const hso_reproj_cand*)(d + o_c), (const uint8_t* const*)(d + o_rp), (AlignJobDev*)(d + o_j),
                      (hso_reproj_result*)(d + o_res), ctx->stream, &ctx->launches));
  tm.align_begin();
  CU(launch_align(ctx->geom, fc->pyr, fc->sobel, (const AlignJobDev*)(d + o_j), M, grid->align_max_iter, (hso_align_result*)(d + o_ar), ctx->stream,
                  &ctx->launches));
  tm.align_end();
  ReprojSelParams sp;
  sp.M = M; sp.n_cells = n_cells; sp.max_fts = grid->max_fts;
  sp.n_sort = 2;
  while (sp.n_sort < M) sp.n_sort *= 2;
  CU(launch_reproj_select(sp, (const hso_reproj_cand*)(d + o_c), (const hso_align_result*)(d + o_ar), (const int32_t*)(d + o_co),
                          (hso_reproj_result*)(d + o_res), (hso_reproj_summary*)(d + o_sum), ctx->stream, &ctx->launches));
  // results and summary are adjacent in the arena only up to alignment: two copies
  char* oh = (char*)ctx->r_out_host.p;
  CU(cudaMemcpyAsync(oh, d + o_res, sizeof(hso_reproj_result) * M, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(oh + sizeof(hso_reproj_result) * M, d + o_sum, sizeof(hso_reproj_summary), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out, oh, sizeof(hso_reproj_result) * M);
  memcpy(summary, oh + sizeof(hso_reproj_result) * M, sizeof(hso_reproj_summary));
  tm.stop_after_sync();
  return HSO_OK;
}

// ---- a13b: seed stage of Reprojector::reprojectMap ---------------------------------------------------------------------------------------------
int hso_reproject_seeds(hso_ctx* ctx, hso_frame_id cur, const double T_cur_w[12], int n_poses, const double* T_f_w, int S, const hso_seed_obs* seeds,
                        const hso_reproj_grid* grid, const int32_t* cell_order, int n_matches_in, hso_reproj_result* out, hso_reproj_summary* summary) {
  if (!ctx || !T_cur_w || !grid || !summary || S < 0 || n_poses < 0 || n_matches_in < 0) return HSO_ERR_INVALID;
  memset(summary, 0, sizeof *summary);
  summary->n_matches = n_matches_in;
  const int n_cells = grid->n_cols * grid->n_rows;
  if (grid->cell_size <= 0 || grid->n_cols <= 0 || grid->n_rows <= 0 || n_cells > 4096 || grid->max_fts < 0 || !cell_order)
    return fail(ctx, HSO_ERR_INVALID, "bad reprojection grid (cells must be <= 4096)");
  if ((long long)grid->n_cols * grid->cell_size < ctx->cam.width || (long long)grid->n_rows * grid->cell_size < ctx->cam.height)
    return fail(ctx, HSO_ERR_INVALID, "reprojection grid does not cover the image");
  if (S == 0) return HSO_OK;
  if (!seeds || !out || !T_f_w || n_poses == 0) return HSO_ERR_INVALID;
  if (S > 16384) return fail(ctx, HSO_ERR_CAPACITY, "more than 16384 seeds");
  FrameSlot* fc = get_frame(ctx, cur);
  if (!fc) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown current frame id");
  {
    std::vector<uint8_t> seen(n_cells, 0);
    for (int i = 0; i < n_cells; ++i) {
      if (cell_order[i] < 0 || cell_order[i] >= n_cells || seen[cell_order[i]]) return fail(ctx, HSO_ERR_INVALID, "cell_order is not a permutation");
      seen[cell_order[i]] = 1;
    }
  }
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = (o + 127) / 128 * 128; o = r + bytes; return r; };
  const size_t o_s = take(sizeof(hso_seed_obs) * S), o_rp = take(sizeof(void*) * S), o_T = take(sizeof(double) * 12 * n_poses),
               o_co = take(sizeof(int32_t) * n_cells);
  const size_t staged = o;
  const size_t o_j = take(sizeof(AlignJobDev) * S), o_ar = take(sizeof(hso_align_result) * S), o_res = take(sizeof(hso_reproj_result) * S),
               o_sum = take(sizeof(hso_reproj_summary));
  CU(ctx->r_arena.reserve(o));
  CU(ctx->r_stage_host.reserve(staged));
  CU(ctx->r_out_host.reserve(sizeof(hso_reproj_result) * S + sizeof(hso_reproj_summary)));
  char* h = (char*)ctx->r_stage_host.p;
  char* d = (char*)ctx->r_arena.p;
  memcpy(h + o_s, seeds, sizeof(hso_seed_obs) * S);
  const uint8_t** rp = (const uint8_t**)(h + o_rp);
  for (int i = 0; i < S; ++i) {
    const hso_seed_obs& sd = seeds[i];
    if (sd.ref_pose < 0 || sd.ref_pose >= n_poses) return fail(ctx, HSO_ERR_INVALID, "pose index out of range in seed");
    FrameSlot* fr = get_frame(ctx, sd.ref_frame);
    if (!fr) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown reference frame id in seed");
    if (sd.level < 0 || sd.level >= ctx->geom.n_levels) return fail(ctx, HSO_ERR_INVALID, "seed level out of range");
    if (sd.ftr_type == 1 && !fc->sobel) return fail(ctx, HSO_ERR_INVALID, "edgelet seeds need hso_cfg.materialize_sobel (checkNormal reads sobelX_/Y_)");
    if (!(sd.sigma2 >= 0.f)) return fail(ctx, HSO_ERR_INVALID, "seed variance must be non-negative");
    rp[i] = fr->pyr;
  }
  memcpy(h + o_T, T_f_w, sizeof(double) * 12 * n_poses);
  memcpy(h + o_co, cell_order, sizeof(int32_t) * n_cells);
  StageTimer tm(ctx, ST_REPROJECT);
  CU(cudaMemcpyAsync(d, h, staged, cudaMemcpyHostToDevice, ctx->stream));
  ReprojKParams kp;
  memset(&kp, 0, sizeof kp);
  kp.cam = ctx->camdev;
  memcpy(kp.T_cur_w, T_cur_w, sizeof kp.T_cur_w);
  kp.T_f_w = (const double*)(d + o_T);
  kp.n_poses = n_poses; kp.M = S; kp.cell_size = grid->cell_size; kp.n_cols = grid->n_cols;
  kp.max_search_level = std::min(ctx->cfg.n_pyr_levels, ctx->geom.n_levels) - 1;
  CU(launch_reproject_seed(kp, (const hso_seed_obs*)(d + o_s), (const uint8_t* const*)(d + o_rp), (AlignJobDev*)(d + o_j), (hso_reproj_result*)(d + o_res),
                           ctx->stream, &ctx->launches));
  tm.align_begin();
  CU(launch_align(ctx->geom, fc->pyr, fc->sobel, (const AlignJobDev*)(d + o_j), S, grid->align_max_iter, (hso_align_result*)(d + o_ar), ctx->stream,
                  &ctx->launches));
  tm.align_end();
  SeedSelParams sp;
  sp.S = S; sp.n_cells = n_cells; sp.max_fts = grid->max_fts; sp.n_matches_in = n_matches_in;
  sp.n_sort = 2;
  while (sp.n_sort < S) sp.n_sort *= 2;
  CU(launch_seed_select(sp, (const hso_seed_obs*)(d + o_s), (const hso_align_result*)(d + o_ar), (const int32_t*)(d + o_co), (hso_reproj_result*)(d + o_res),
                        (hso_reproj_summary*)(d + o_sum), ctx->stream, &ctx->launches));
  char* oh = (char*)ctx->r_out_host.p;
  CU(cudaMemcpyAsync(oh, d + o_res, sizeof(hso_reproj_result) * S, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(oh + sizeof(hso_reproj_result) * S, d + o_sum, sizeof(hso_reproj_summary), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out, oh, sizeof(hso_reproj_result) * S);
  memcpy(summary, oh + sizeof(hso_reproj_result) * S, sizeof(hso_reproj_summary));
  tm.stop_after_sync();
  return HSO_OK;
}

// Test hook for row N1: only the selection kernel, on caller-given per-candidate facts (what k_reproject / k_align would have produced).
int hso_reproject_select_only(hso_ctx* ctx, int M, const hso_reproj_cand* cands, const int32_t* in_frame, const int32_t* cell, const uint8_t* align_ok,
                              const hso_reproj_grid* grid, const int32_t* cell_order, hso_reproj_result* out, hso_reproj_summary* summary) {
  if (!ctx || M <= 0 || M > 16384 || !cands || !in_frame || !cell || !align_ok || !grid || !cell_order || !out || !summary) return HSO_ERR_INVALID;
  const int n_cells = grid->n_cols * grid->n_rows;
  if (n_cells <= 0 || n_cells > 4096) return fail(ctx, HSO_ERR_INVALID, "bad reprojection grid (cells must be <= 4096)");
  for (int i = 0; i < M; ++i)
    if (in_frame[i] && (cell[i] < 0 || cell[i] >= n_cells)) return fail(ctx, HSO_ERR_INVALID, "cell index out of range");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = (o + 127) / 128 * 128; o = r + bytes; return r; };
  const size_t o_c = take(sizeof(hso_reproj_cand) * M), o_ar = take(sizeof(hso_align_result) * M), o_res = take(sizeof(hso_reproj_result) * M),
               o_co = take(sizeof(int32_t) * n_cells);
  const size_t staged = o;
  const size_t o_sum = take(sizeof(hso_reproj_summary));
  CU(ctx->r_arena.reserve(o));
  CU(ctx->r_stage_host.reserve(staged));
  CU(ctx->r_out_host.reserve(sizeof(hso_reproj_result) * M + sizeof(hso_reproj_summary)));
  char* h = (char*)ctx->r_stage_host.p;
  char* d = (char*)ctx->r_arena.p;
  memcpy(h + o_c, cands, sizeof(hso_reproj_cand) * M);
  hso_align_result* har = (hso_align_result*)(h + o_ar);
  hso_reproj_result* hres = (hso_reproj_result*)(h + o_res);
  memset(har, 0, sizeof(hso_align_result) * M);
  memset(hres, 0, sizeof(hso_reproj_result) * M);
  for (int i = 0; i < M; ++i) {
    har[i].ok = align_ok[i] ? 1 : 0;
    hres[i].in_frame = in_frame[i] ? 1 : 0; hres[i].cell = in_frame[i] ? cell[i] : -1; hres[i].order = -1;
  }
  memcpy(h + o_co, cell_order, sizeof(int32_t) * n_cells);
  CU(cudaMemcpyAsync(d, h, staged, cudaMemcpyHostToDevice, ctx->stream));
  ReprojSelParams sp;
  sp.M = M; sp.n_cells = n_cells; sp.max_fts = grid->max_fts;
  sp.n_sort = 2;
  while (sp.n_sort < M) sp.n_sort *= 2;
  CU(launch_reproj_select(sp, (const hso_reproj_cand*)(d + o_c), (const hso_align_result*)(d + o_ar), (const int32_t*)(d + o_co),
                          (hso_reproj_result*)(d + o_res), (hso_reproj_summary*)(d + o_sum), ctx->stream, &ctx->launches));
  char* oh = (char*)ctx->r_out_host.p;
  CU(cudaMemcpyAsync(oh, d + o_res, sizeof(hso_reproj_result) * M, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(oh + sizeof(hso_reproj_result) * M, d + o_sum, sizeof(hso_reproj_summary), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out, oh, sizeof(hso_reproj_result) * M);
  memcpy(summary, oh + sizeof(hso_reproj_result) * M, sizeof(hso_reproj_summary));
  return HSO_OK;
}

// ---- N3 ---------------------------------------------------------------------------------------------------------------------
int hso_depth_observe(hso_ctx* ctx, hso_frame_id cur, const double T_cur_w[12], int n_poses, const double* T_f_w, double px_error_angle,
                      int align_max_iter, int S, const hso_seed_obs* seeds, hso_seed_result* out) {
  if (!ctx || !T_cur_w || S < 0 || n_poses < 0 || align_max_iter < 0) return HSO_ERR_INVALID;
  if (S == 0) return HSO_OK;
  if (!seeds || !out || !T_f_w || n_poses == 0) return HSO_ERR_INVALID;
  FrameSlot* fc = get_frame(ctx, cur);
  if (!fc) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown active frame id");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = (o + 127) / 128 * 128; o = r + bytes; return r; };
  const size_t o_s = take(sizeof(hso_seed_obs) * S), o_rp = take(sizeof(void*) * S), o_T = take(sizeof(double) * 12 * n_poses);
  const size_t staged = o;
  const size_t o_out = take(sizeof(hso_seed_result) * S);
  CU(ctx->d_arena.reserve(o));
  CU(ctx->d_stage_host.reserve(staged));
  CU(ctx->d_out_host.reserve(sizeof(hso_seed_result) * S));
  char* h = (char*)ctx->d_stage_host.p;
  char* d = (char*)ctx->d_arena.p;
  memcpy(h + o_s, seeds, sizeof(hso_seed_obs) * S);
  const uint8_t** rp = (const uint8_t**)(h + o_rp);
  for (int i = 0; i < S; ++i) {
    const hso_seed_obs& s = seeds[i];
    if (s.ref_pose < 0 || s.ref_pose >= n_poses) return fail(ctx, HSO_ERR_INVALID, "pose index out of range in seed");
    FrameSlot* fr = get_frame(ctx, s.ref_frame);
    if (!fr) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown reference frame id in seed");
    if (s.level < 0 || s.level >= ctx->geom.n_levels) return fail(ctx, HSO_ERR_INVALID, "seed level out of range");
    if (s.ftr_type == 1 && !fc->sobel) return fail(ctx, HSO_ERR_INVALID, "edgelet seeds need hso_cfg.materialize_sobel (checkNormal reads sobelX_/Y_)");
    rp[i] = fr->pyr;
  }
  memcpy(h + o_T, T_f_w, sizeof(double) * 12 * n_poses);
  StageTimer tm(ctx, ST_DEPTH_FILTER);
  CU(cudaMemcpyAsync(d, h, staged, cudaMemcpyHostToDevice, ctx->stream));
  DepthKParams kp;
  memset(&kp, 0, sizeof kp);
  kp.g = ctx->geom; kp.cam = ctx->camdev;
  memcpy(kp.T_cur_w, T_cur_w, sizeof kp.T_cur_w);
  kp.T_f_w = (const double*)(d + o_T);
  kp.S = S; kp.max_iter = align_max_iter;
  kp.max_search_level = std::min(ctx->cfg.n_pyr_levels, ctx->geom.n_levels) - 1;
  kp.px_error_angle = px_error_angle;
  kp.cur_pyr = fc->pyr; kp.cur_sobel = fc->sobel;
  size_t so = 0;
  for (int l = 0; l < 3; ++l) {
    kp.sobel_off[l] = so;
    if (l < ctx->geom.n_levels) so += (size_t)2 * ctx->geom.w[l] * ctx->geom.h[l];
  }
  CU(launch_depth_observe(kp, (const hso_seed_obs*)(d + o_s), (const uint8_t* const*)(d + o_rp), (hso_seed_result*)(d + o_out), ctx->stream,
                          &ctx->launches));
  CU(cudaMemcpyAsync(ctx->d_out_host.p, d + o_out, sizeof(hso_seed_result) * S, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->d_out_host.p, sizeof(hso_seed_result) * S);
  tm.stop_after_sync();
  return HSO_OK;
}

// ---- F4 ---------------------------------------------------------------------------------------------------------------------
int hso_pose_optimize_batch(hso_ctx* ctx, double reproj_thresh, int n_iter, int B, const int32_t* n_fts_total, const int32_t* offs,
                            const double* f, const double* p_host, const int32_t* host_idx, const int32_t* hoffs, const double* T_host_w,
                            const double* grad, const int8_t* level, const int8_t* ftype, const int8_t* ptype, const double* T_f_w_in,
                            uint8_t* outlier_out, hso_pose_result* out) {
  if (!ctx || B <= 0 || !n_fts_total || !offs || !hoffs || !T_f_w_in || !out || n_iter < 0) return HSO_ERR_INVALID;
  const int Ftot = offs[B], Ktot = hoffs[B];
  if (Ftot < 0 || Ktot < 0) return HSO_ERR_INVALID;
  if (Ftot > 0 && (!f || !p_host || !host_idx || !T_host_w || !grad || !level || !ftype || !ptype || !outlier_out)) return HSO_ERR_INVALID;
  for (int b = 0; b < B; ++b) {
    const int F = offs[b + 1] - offs[b], K = hoffs[b + 1] - hoffs[b];
    if (F < 0 || K < 0) return HSO_ERR_INVALID;
    for (int i = offs[b]; i < offs[b + 1]; ++i)
      if (host_idx[i] < 0 || host_idx[i] >= K) return fail(ctx, HSO_ERR_INVALID, "host_idx out of range");
  }
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  // one staging blob: [f | p_host | grad | T_host_w | host_idx | level | ftype | ptype] then device-only scratch
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = (o + 63) / 64 * 64; o = r + bytes; return r; };
  const size_t o_f = take(sizeof(double) * 3 * Ftot), o_ph = take(sizeof(double) * 3 * Ftot), o_g = take(sizeof(double) * 2 * Ftot);
  const size_t o_th = take(sizeof(double) * 12 * Ktot), o_hi = take(sizeof(int32_t) * Ftot);
  const size_t o_lv = take(Ftot), o_ft = take(Ftot), o_pt = take(Ftot);
  const size_t o_jobs = take(sizeof(PoseJobDev) * B), o_scr = take(sizeof(PoseScratch) * B);
  const size_t staged = o;
  const size_t o_inv = take(sizeof(Se3d) * Ktot), o_tth = take(sizeof(double) * 12 * Ktot), o_keys = take(sizeof(unsigned long long) * Ftot);
  const size_t o_cls = take(Ftot), o_outl = take(Ftot), o_res = take(sizeof(hso_pose_result) * B);
  CU(ctx->p_arena.reserve(o));
  CU(ctx->p_stage_host.reserve(staged));
  CU(ctx->p_out_host.reserve(sizeof(hso_pose_result) * B + Ftot));
  char* h = (char*)ctx->p_stage_host.p;
  char* d = (char*)ctx->p_arena.p;
  if (Ftot > 0) {
    memcpy(h + o_f, f, sizeof(double) * 3 * Ftot); memcpy(h + o_ph, p_host, sizeof(double) * 3 * Ftot);
    memcpy(h + o_g, grad, sizeof(double) * 2 * Ftot); memcpy(h + o_hi, host_idx, sizeof(int32_t) * Ftot);
    memcpy(h + o_lv, level, Ftot); memcpy(h + o_ft, ftype, Ftot); memcpy(h + o_pt, ptype, Ftot);
  }
  if (Ktot > 0) memcpy(h + o_th, T_host_w, sizeof(double) * 12 * Ktot);
  PoseJobDev* hj = (PoseJobDev*)(h + o_jobs);
  PoseScratch* hs = (PoseScratch*)(h + o_scr);
  for (int b = 0; b < B; ++b) {
    const int F0 = offs[b], K0 = hoffs[b];
    PoseJobDev& j = hj[b];
    j.F = offs[b + 1] - F0; j.K = hoffs[b + 1] - K0; j.n_fts_total = n_fts_total[b]; j.pad_ = 0;
    j.f = (const double*)(d + o_f) + 3 * F0; j.p_host = (const double*)(d + o_ph) + 3 * F0;
    j.host_idx = (const int32_t*)(d + o_hi) + F0; j.T_host_w = (const double*)(d + o_th) + 12 * K0;
    j.grad = (const double*)(d + o_g) + 2 * F0;
    j.level = (const int8_t*)(d + o_lv) + F0; j.ftype = (const int8_t*)(d + o_ft) + F0; j.ptype = (const int8_t*)(d + o_pt) + F0;
    memcpy(j.T_f_w_in, T_f_w_in + 12 * b, sizeof(double) * 12);
    j.outlier = (uint8_t*)(d + o_outl) + F0;
    j.scratch = nullptr;
    j.out = (hso_pose_result*)(d + o_res) + b;
    hs[b].T_host_inv = (Se3d*)(d + o_inv) + K0; hs[b].Tth = (double*)(d + o_tth) + 12 * K0;
    hs[b].keys = (unsigned long long*)(d + o_keys) + F0; hs[b].cls = (int8_t*)(d + o_cls) + F0;
  }
  // AbstractCamera::errorMultiplier2() = fxy_mean_ (src/camera.cpp:59,161,287)
  const double fx = ctx->cam.fx, fy = ctx->cam.fy;
  const double err_mult2 = (fx * fy < 0) ? std::fabs(fx) : std::fabs((fx + fy) * 0.5);
  StageTimer tm(ctx, 3);
  CU(cudaMemcpyAsync(d, h, staged, cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_pose((const PoseJobDev*)(d + o_jobs), (const PoseScratch*)(d + o_scr), B, reproj_thresh, n_iter, err_mult2, ctx->stream,
                 &ctx->launches));
  char* ho = (char*)ctx->p_out_host.p;
  CU(cudaMemcpyAsync(ho, d + o_res, sizeof(hso_pose_result) * B, cudaMemcpyDeviceToHost, ctx->stream));
  if (Ftot > 0) CU(cudaMemcpyAsync(ho + sizeof(hso_pose_result) * B, d + o_outl, Ftot, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ho, sizeof(hso_pose_result) * B);
  if (Ftot > 0) memcpy(outlier_out, ho + sizeof(hso_pose_result) * B, Ftot);
  tm.stop_after_sync();
  return HSO_OK;
}

int hso_pose_optimize(hso_ctx* ctx, double reproj_thresh, int n_iter, int n_fts_total, int F, const double* f, const double* p_host,
                      const int32_t* host_idx, int K, const double* T_host_w, const double* grad, const int8_t* level, const int8_t* ftype,
                      const int8_t* ptype, const double T_f_w_in[12], uint8_t* outlier_out, hso_pose_result* out) {
  const int32_t offs[2] = {0, F}, hoffs[2] = {0, K}, nf[1] = {n_fts_total};
  return hso_pose_optimize_batch(ctx, reproj_thresh, n_iter, 1, nf, offs, f, p_host, host_idx, hoffs, T_host_w, grad, level, ftype, ptype,
                                 T_f_w_in, outlier_out, out);
}

// ---- N2 -------------------------------------------------------------------------------------------------------------------------
int hso_fast_detect(hso_ctx* ctx, hso_frame_id frame, int level, int threshold, int border, hso_corner* out, int cap, int* count) {
  if (!ctx || level < 0 || level >= ctx->geom.n_levels || cap < 0 || (cap > 0 && !out) || !count || threshold < 0 || threshold > 254)
    return HSO_ERR_INVALID;
  FrameSlot* f = get_frame(ctx, frame);
  if (!f) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown frame id");
  const int w = ctx->geom.w[level], h = ctx->geom.h[level];
  *count = 0;
  if (w < 7 || h < 7) return HSO_OK;  // fast_corner_detect_9_sse2 returns nothing (faster_corner_9_sse.cpp:246-250)
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  StageTimer tm(ctx, ST_FEATURE_DETECT);
  const size_t npx = (size_t)ctx->geom.w[0] * ctx->geom.h[0];
  CU(ctx->f_score.reserve(npx * sizeof(int16_t)));
  CU(ctx->f_rowbuf.reserve(npx * sizeof(uint32_t)));
  CU(ctx->f_rowcount.reserve(sizeof(int) * ctx->geom.h[0]));
  CU(ctx->f_total.reserve(sizeof(int)));
  CU(ctx->f_out.reserve(sizeof(hso_corner) * (size_t)std::max(cap, 1)));
  CU(ctx->f_out_host.reserve(sizeof(hso_corner) * (size_t)std::max(cap, 1) + 4 * sizeof(int)));
  CU(launch_fast(f->pyr + ctx->geom.off[level], w, h, threshold, border, (int16_t*)ctx->f_score.p, (uint32_t*)ctx->f_rowbuf.p, (int*)ctx->f_rowcount.p,
                 (hso_corner*)ctx->f_out.p, cap, (int*)ctx->f_total.p, ctx->stream, &ctx->launches));
  int* total_h = (int*)ctx->f_out_host.p;
  hso_corner* out_h = (hso_corner*)(total_h + 4);
  CU(cudaMemcpyAsync(total_h, ctx->f_total.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const int n = std::min(*total_h, cap);
  if (n > 0) {
    CU(cudaMemcpyAsync(out_h, ctx->f_out.p, sizeof(hso_corner) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(out, out_h, sizeof(hso_corner) * n);
  }
  *count = *total_h;
  tm.stop_after_sync();
  return HSO_OK;
}

// FeatureExtractor::fastDetectMT runs the three pyramid levels on three host threads (src/feature_detection.cpp:498-514); here the nine kernels of
// levels 0 .. n_levels-1 go out back to back and the corner lists come back with ONE synchronisation in the common case (the totals and the
// first kSpec corners of every level travel together; a second copy fetches the rest of a level that found more).
int hso_fast_detect_levels(hso_ctx* ctx, hso_frame_id frame, int n_levels, int threshold, int border, hso_corner* out, int cap_per_level, int* counts) {
  if (!ctx || n_levels <= 0 || n_levels > 3 || n_levels > ctx->geom.n_levels || cap_per_level < 0 || (cap_per_level > 0 && !out) || !counts ||
      threshold < 0 || threshold > 254)
    return HSO_ERR_INVALID;
  FrameSlot* f = get_frame(ctx, frame);
  if (!f) return fail(ctx, HSO_ERR_BAD_FRAME, "unknown frame id");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  StageTimer tm(ctx, ST_FEATURE_DETECT);
  const int kSpec = std::min(cap_per_level, 4096);
  const size_t npx = (size_t)ctx->geom.w[0] * ctx->geom.h[0];
  CU(ctx->f_score.reserve(npx * sizeof(int16_t)));
  CU(ctx->f_rowbuf.reserve(npx * sizeof(uint32_t)));
  CU(ctx->f_rowcount.reserve(sizeof(int) * ctx->geom.h[0]));
  CU(ctx->f_total.reserve(sizeof(int) * 4));
  CU(ctx->f_out.reserve(sizeof(hso_corner) * (size_t)std::max(cap_per_level, 1) * n_levels));
  CU(ctx->f_out_host.reserve(sizeof(hso_corner) * (size_t)std::max(cap_per_level, 1) * n_levels + 4 * sizeof(int)));
  CU(cudaMemsetAsync(ctx->f_total.p, 0, sizeof(int) * 4, ctx->stream));
  int* total_h = (int*)ctx->f_out_host.p;
  hso_corner* out_h = (hso_corner*)(total_h + 4);
  for (int l = 0; l < n_levels; ++l) {
    const int w = ctx->geom.w[l], h = ctx->geom.h[l];
    counts[l] = 0;
    if (w < 7 || h < 7) continue;  // fast_corner_detect_9_sse2 returns nothing (faster_corner_9_sse.cpp:246-250)
    CU(launch_fast(f->pyr + ctx->geom.off[l], w, h, threshold, border, (int16_t*)ctx->f_score.p, (uint32_t*)ctx->f_rowbuf.p, (int*)ctx->f_rowcount.p,
                   (hso_corner*)ctx->f_out.p + (size_t)l * cap_per_level, cap_per_level, (int*)ctx->f_total.p + l, ctx->stream, &ctx->launches));
    if (kSpec > 0)
      CU(cudaMemcpyAsync(out_h + (size_t)l * cap_per_level, (hso_corner*)ctx->f_out.p + (size_t)l * cap_per_level, sizeof(hso_corner) * kSpec,
                         cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaMemcpyAsync(total_h, ctx->f_total.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  bool more = false;
  for (int l = 0; l < n_levels; ++l) {
    counts[l] = total_h[l];
    const int n = std::min(total_h[l], cap_per_level);
    if (n > kSpec) {
      more = true;
      CU(cudaMemcpyAsync(out_h + (size_t)l * cap_per_level + kSpec, (hso_corner*)ctx->f_out.p + (size_t)l * cap_per_level + kSpec,
                         sizeof(hso_corner) * (n - kSpec), cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  if (more) CU(cudaStreamSynchronize(ctx->stream));
  for (int l = 0; l < n_levels; ++l) {
    const int n = std::min(total_h[l], cap_per_level);
    if (n > 0) memcpy(out + (size_t)l * cap_per_level, out_h + (size_t)l * cap_per_level, sizeof(hso_corner) * n);
  }
  tm.stop_after_sync();
  return HSO_OK;
}

const char* hso_stage_name(int stage) { return (stage >= 0 && stage < kStages) ? kStageNames[stage] : nullptr; }

int hso_stage_time_ms(hso_ctx* ctx, int stage, double* ms, uint64_t* calls) {
  if (!ctx || stage < 0 || stage >= kStages) return HSO_ERR_INVALID;
  if (ms) *ms = ctx->stage_ms[stage];
  if (calls) *calls = ctx->stage_calls[stage];
  return HSO_OK;
}

}  // extern "C"
