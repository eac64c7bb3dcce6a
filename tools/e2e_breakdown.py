"""Where does an e2e step go? Times the C-ABI calls of bench.py's e2e step one by one (host wall clock, each call is synchronous)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from hso_b200 import Context, make_cam, _capi as K

B, F = int(os.environ.get("HSO_BD_BATCH", "1184")), 3000
probs = bench.build_workload(B, F, "icl", 0x450, 0)
c = probs[0]["cam"]; W, H = c["width"], c["height"]
ctx = Context(make_cam(W, H, c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), device=0, max_frames=2 * B + 72, max_features=8192)
lib = ctx.lib
host_cur = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
for b, p in enumerate(probs):
    host_cur[b].copy_(torch.from_numpy(p["cur_img"]))
cur_np = [host_cur[b].numpy() for b in range(B)]
ref_ids, ref_int, _ = ctx.upload_frames([p["ref_img"] for p in probs])
cur_ids, cur_int, _ = ctx.upload_frames(cur_np)
jobs = [dict(ref=ref_ids[b], cur=cur_ids[b], px=p["px"], f=p["f"], dist=p["dist"], T_cur_ref=p["T0"], exposure_rat=1.0) for b, p in enumerate(probs)]
prm = K.hso_track_params(0, 4, 1, 50)
jarr, keep = ctx._track_jobs(jobs)
img_ptrs = (C.c_void_p * B)(*[im.ctypes.data for im in cur_np])
new_ids = (C.c_int32 * B)(); integ = np.zeros(B, np.float32); res = (K.hso_track_result * B)()
cur_ids_c = (C.c_int32 * B)(*cur_ids)
fptr = C.POINTER(C.c_float)
# raw pinned H2D bandwidth
dst = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); dst.copy_(host_cur, non_blocking=True); torch.cuda.synchronize()
    print(f"raw pinned H2D {host_cur.numel()/1e6:.0f} MB: {1e3*(time.perf_counter()-t):.2f} ms = {host_cur.numel()/1e9/(time.perf_counter()-t):.1f} GB/s")
T = {}
def tick(name, t0):
    T.setdefault(name, []).append(1e3 * (time.perf_counter() - t0))
for it in range(6):
    t = time.perf_counter()
    for b in range(B): lib.hso_frame_release(ctx.h, cur_ids_c[b])
    tick("release(py loop)", t); t = time.perf_counter()
    ctx._chk(lib.hso_frame_upload_batch(ctx.h, B, img_ptrs, W, H, W, new_ids, integ.ctypes.data_as(fptr), None))
    tick("frame_upload_batch", t); t = time.perf_counter()
    for b in range(B):
        cur_ids_c[b] = new_ids[b]; jarr[b].cur = new_ids[b]; jarr[b].exposure_rat = 1.0
    tick("py job refresh", t); t = time.perf_counter()
    ctx._chk(lib.hso_track_stage(ctx.h, C.byref(prm), B, jarr, 0))
    tick("track_stage(host)", t); t = time.perf_counter()
    ctx._chk(lib.hso_synchronize(ctx.h))
    tick("track_stage(H2D wait)", t); t = time.perf_counter()
    ctx._chk(lib.hso_track_run(ctx.h)); ctx._chk(lib.hso_synchronize(ctx.h))
    tick("track_run", t); t = time.perf_counter()
    ctx._chk(lib.hso_track_collect(ctx.h, res, None, None))
    tick("track_collect", t)
for k, v in T.items():
    print(f"{k:26s} {np.median(v[2:]):8.3f} ms")
print("total", sum(np.median(v[2:]) for v in T.values()))
# ---- the chunk-pipelined single call ----------------------------------------------------------------------------------------
gm = np.zeros(B, np.float32)
for b in range(B):
    jarr[b].exposure_rat = -1.0
for chunk, streams in [(74, 3), (74, 4), (56, 4), (111, 3), (111, 4), (148, 4), (0, 0)]:
    ctx._chk(lib.hso_set_pipeline(ctx.h, chunk, streams))
    tt = []
    for it in range(8):
        for b in range(B): lib.hso_frame_release(ctx.h, cur_ids_c[b])
        t = time.perf_counter()
        ctx._chk(lib.hso_add_frames_track_batch(ctx.h, C.byref(prm), B, img_ptrs, W, H, W, jarr, new_ids, integ.ctypes.data_as(fptr), gm.ctypes.data_as(fptr), res))
        tt.append(1e3 * (time.perf_counter() - t))
        for b in range(B): cur_ids_c[b] = new_ids[b]
    print(f"hso_add_frames_track_batch chunk={chunk} streams={streams} ms: median {np.median(tt[2:]):.2f} min {min(tt):.2f}  iters", sum(res[b].n_iters for b in range(B)))

# ---- the stages of the pipelined call alone (HSO_PIPE_DEBUG): copies + host flattening only / kernels only (data resident from the run above)
ctx._chk(lib.hso_set_pipeline(ctx.h, 0, 0))
for name, flag in (("copies + flattening only", "1"), ("image copies + flattening only", "9"), ("feature copies + flattening only", "5"), ("flattening only", "3"), ("kernels only (no image / feature copies)", "2"), ("everything", "0")):
    os.environ["HSO_PIPE_DEBUG"] = flag
    tt = []
    for it in range(6):
        for b in range(B): lib.hso_frame_release(ctx.h, cur_ids_c[b])
        t = time.perf_counter()
        ctx._chk(lib.hso_add_frames_track_batch(ctx.h, C.byref(prm), B, img_ptrs, W, H, W, jarr, new_ids, integ.ctypes.data_as(fptr), gm.ctypes.data_as(fptr), res))
        tt.append(1e3 * (time.perf_counter() - t))
        for b in range(B): cur_ids_c[b] = new_ids[b]
    print(f"pipelined call, {name}: median {np.median(tt[2:]):.2f} ms")
# ---- kernels-only / everything per pipeline shape
for chunk, streams in [(111, 3), (148, 3), (148, 4), (222, 3), (222, 4), (296, 3), (296, 4), (74, 4)]:
    ctx._chk(lib.hso_set_pipeline(ctx.h, chunk, streams))
    line = f"chunk={chunk} streams={streams}:"
    for name, flag in (("kernels", "2"), ("all", "0")):
        os.environ["HSO_PIPE_DEBUG"] = flag
        tt = []
        for it in range(6):
            for b in range(B): lib.hso_frame_release(ctx.h, cur_ids_c[b])
            t = time.perf_counter()
            ctx._chk(lib.hso_add_frames_track_batch(ctx.h, C.byref(prm), B, img_ptrs, W, H, W, jarr, new_ids, integ.ctypes.data_as(fptr), gm.ctypes.data_as(fptr), res))
            tt.append(1e3 * (time.perf_counter() - t))
            for b in range(B): cur_ids_c[b] = new_ids[b]
        line += f"  {name} {np.median(tt[2:]):.2f} ms"
    print(line)
