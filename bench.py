#!/usr/bin/env python
"""bench.py — throughput of the tracking hot path (BASELINE.json metric: CoarseTracker LM iterations/s @640x480, 3000 patches).

A step = one pass of the hot path over one batch of B synthetic frame pairs on one GPU: Frame construction (pyramid + gradient
statistics) of the B current images, then CoarseTracker::run L4->L1 (n_iter = 50, natural convergence) of every current frame
against its reference frame. `value` = LM iterations (trials) executed by the whole job per second, inputs resident in HBM.
`e2e` = the same metric through the C-ABI with HOST buffers (hso_add_frames_track_batch: H2D of the images and feature arrays,
D2H of the results inside the timed region; the call pipelines copies against kernels chunk by chunk).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Multi-GPU: independent VO streams are sharded one batch per GPU (weak scaling); NCCL only for the barrier and the max-over-ranks /
sum-over-ranks of three counters. --impl reference times the CPU oracle (the reference algorithm restated in C, oracle/) on all
host threads on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hso_b200 import synth  # noqa: E402

# algorithmic bytes per visible patch per residual evaluation, forward mode, by level
BYTES_PER_PATCH_EVAL = {4: 88, 3: 132, 2: 132, 1: 200, 0: 180}
BYTES_PER_PATCH_EVAL_IC = {4: 188, 3: 256, 2: 256, 1: 380, 0: 400}
N_BASE = 12  # distinct synthetic scenes; problems cycle through them with different features / initial poses
TRACK_CFG = dict(min_level=1, n_iter=50)  # set from --min-level / --n-iter; the CPU legs run the same settings


def build_workload(B, F, cam, seed0, rank):
    """B (ref, cur) problems at the workload's size. Scenes repeat every N_BASE problems (host generation cost), but every
    problem has its own device buffers, its own feature subset and its own initial pose."""
    rng = np.random.default_rng(seed0 + 7919 * rank)
    bases = [synth.make_pair(seed0 + 100 * rank + i, cam, F=F, motion_scale=1.0 + 0.5 * (i % 3)) for i in range(min(N_BASE, B))]
    probs = []
    for b in range(B):
        base = bases[b % len(bases)]
        c = base["cam"]
        W, H = c["width"], c["height"]
        px = np.stack([rng.uniform(8, W - 8, F), rng.uniform(8, H - 8, F)], axis=1)
        rx, ry = synth.cam2world_plane(c, px[:, 0], px[:, 1])  # bearings through the camera model (Feature::f = cam2world(px))
        ray = np.stack([rx, ry, np.ones(F)], axis=1)
        f = ray / np.linalg.norm(ray, axis=1, keepdims=True)
        dist = 4.0 / f[:, 2]
        dist[rng.uniform(size=F) < 0.02] = -1.0
        if c.get("model", 0) == 1 and not c.get("undistort", 0):  # pixels outside the FOV model's domain carry no point
            dist[np.hypot((px[:, 0] - c["cx"]) / c["fx"], (px[:, 1] - c["cy"]) / c["fy"]) * c["d"][0] > 1.45] = -1.0
        T0 = synth.se3_exp(np.concatenate([rng.normal(0, 0.004, 3), rng.normal(0, 0.001, 3)]))[:3]
        probs.append(dict(base=b % len(bases), ref_img=base["ref_img"], cur_img=base["cur_img"], px=px, f=f, dist=dist, T0=T0, cam=c))
    return probs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        """Samples that arrived inside [t0, t1] (the timed region, host clock); the sampler is started before the warm-up steps because
        nvidia-smi needs ~0.1 s to deliver its first line and the timed region may be shorter than that."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [ln for (ts, ln) in self.lines if t0 is None or (t0 <= ts <= t1 + 0.03)]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: fall back to the samples under the identical warm-up load
            inside, window = [ln for (_, ln) in self.lines], "warm-up + timed region"
        for ln in inside:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


_JSON_OUT = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout. Libraries print there too (NCCL's version banner under torchrun), so the real stdout is
    kept aside for the JSON line and file descriptor 1 is pointed at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def bind_to_gpu_numa_node(gpu_index):
    """Multi-GPU runs: keep this rank's threads (and, by first touch, its pinned staging memory) on the NUMA node its GPU hangs off, so that
    H2D traffic does not cross the socket interconnect. Best effort: any missing sysfs / nvidia-smi piece leaves the affinity untouched."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)], capture_output=True, text=True,
                             timeout=10).stdout.strip().lower()
        if not bus:
            return None
        bus = bus[-12:] if len(bus) > 12 else bus  # sysfs uses a 4-digit domain: 0000:18:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 2:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------------------
def cpu_frame(O, prob, ic):
    """The reference algorithm (oracle PORT) for one frame of the step: pyramid + Sobel + stats of the current image, CoarseTracker::run."""
    cl, _ = O.create_pyramid(prob["cur_img"], 5)
    for l in range(3):
        O.sobel5(cl[l])
    ci, _ = O.frame_stats(prob["cur_img"])
    tp = O.TrackProblem(prob["cam"], prob["_ref_levels"], cl, prob["px"], prob["f"], prob["dist"])
    r = tp.run(prob["T0"], float(np.float32(ci) / np.float32(prob["_ref_integral"])), inverse_comp=ic, trace_cap=64, **TRACK_CFG)
    return r["n_evals"] - (5 - TRACK_CFG["min_level"])  # evaluations minus one entry evaluation per level = LM trials


def ref_frame(R, prob, ic):
    """The REFERENCE ITSELF (oracle/_ref/libhso_ref.so: the reference's sources compiled unmodified) for one frame of the step: new hso::Frame
    (pyramid, Sobel images, statistics: src/frame.cpp:45-96) then CoarseTracker::run(ref, cur) (src/CoarseTracker.cpp:51-208). The reference
    does not report its trial count; the step's LM iterations are counted by the oracle port on the same frames outside the timed region."""
    cur = R.Frame(prob["cam"], prob["cur_img"])
    R.coarse_track(prob["_ref_frame"], cur, prob["T0"], inverse_comp=ic, **TRACK_CFG)
    cur.close()
    return 0


def cpu_prepare(O, probs):
    cache = {}
    for p in probs:
        if p["base"] not in cache:
            rl, _ = O.create_pyramid(p["ref_img"], 5)
            cache[p["base"]] = (rl, O.frame_stats(p["ref_img"])[0])
        p["_ref_levels"], p["_ref_integral"] = cache[p["base"]]


def run_cpu(probs, ic, threads, kind="auto"):
    """Times the CPU implementation of the step on `probs`. kind: "reference" = the reference's own code (oracle/_ref), "port" = the oracle
    restatement, "auto" = reference when its library is present. Returns (LM iterations, seconds, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import ref_lib as R
    O.load()
    if kind == "auto":
        kind = "reference" if R.available() else "port"
    cpu_prepare(O, probs)
    if kind == "reference":
        R.load()
        # every problem gets its own reference Frame object carrying its own features (the reference keeps features inside the Frame);
        # built outside the timed region like the resident reference frames of the GPU arm
        for p in probs:
            p["_ref_frame"] = R.Frame(p["cam"], p["ref_img"])
            p["_ref_frame"].set_track_features(p["px"], p["f"], p["dist"])
        work = lambda p: ref_frame(R, p, ic)
    else:
        work = lambda p: cpu_frame(O, p, ic)
    t0 = time.perf_counter()
    if threads <= 1:
        iters = sum(work(p) for p in probs)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the native code
            iters = sum(ex.map(work, probs))
    dt = time.perf_counter() - t0
    if kind == "reference":
        for p in probs:
            p["_ref_frame"].close()
            del p["_ref_frame"]
        for p in probs:  # LM iterations of these frames, counted once by the port (same algorithm, same trial sequence)
            if "_iters_port" not in p:
                p["_iters_port"] = cpu_frame(O, p, ic)
        iters = sum(p["_iters_port"] for p in probs)
    return iters, dt, kind


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = 4 * max(cores, 2)
    probs = build_workload(n, args.patches, args.cam, args.seed, 0)
    for _ in range(args.warmup):  # W untimed warm-up steps on a small slice (page-in, thread start-up)
        run_cpu(probs[: max(2, cores // 4)], args.ic, cores)
    tot_it, tot_t = 0, 0.0
    steps = max(1, args.steps)  # exactly K timed steps; a step is a bounded sample (4 frames per host thread, ~0.2 s)
    kind = "port"
    for _ in range(steps):
        it, dt, kind = run_cpu(probs, args.ic, cores)
        tot_it += it; tot_t += dt
    what = ("the reference's own sources compiled unmodified (oracle/_ref/libhso_ref.so: new hso::Frame + CoarseTracker::run; cv::Sobel / cv::resize "
            "are the shim's scalar restatements, ~6 % of a frame)" if kind == "reference" else "oracle/liboracle_hso.so (-O3 x86-64-v3)")
    v = tot_it / tot_t
    line = {"impl": "reference", "metric": "CoarseTracker LM iterations/sec @640x480, 3k patches", "value": v, "unit": "iterations/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "frames_per_s": steps * n / tot_t,
            "config": {"workload": f"{args.cam} {probs[0]['cam']['width']}x{probs[0]['cam']['height']}, {args.patches} patches/frame, "
                                   f"pyramid+stats then CoarseTracker L4->L{args.min_level} n_iter={args.n_iter} {'inverse-compositional' if args.ic else 'forward'}",
                       "sample": f"{n} frames per step"},
            "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": cores, "kind": kind,
                             "sample": f"{n} frames/step x {steps} steps, one frame per thread, {what}"},
            "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)



def other_rows(ctx, lib, args, dev, torch, K):
    """Throughput of the other rows of the hot path through the C-ABI with host buffers, beside the CPU oracle:
    Frame construction, direct patch matching (BASELINE config 4: 5000 candidates) and the pose optimiser (5000 features)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    out = {}
    pair = synth.make_pair(args.seed + 5, args.cam, F=8)
    c = pair["cam"]
    W, H = c["width"], c["height"]

    def timed(fn, reps):
        fn()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ctx.synchronize()
        return (time.perf_counter() - t0) / reps

    peak, _src = measured_peak()

    class Roof:
        """Device time of a stage (CUDA events inside the library, hso_stage_time_ms) over a block of calls -> achieved algorithmic GB/s
        against the measured HBM peak. These rows run one small problem per call: they are launch / latency bound, the fraction says so."""

        def __init__(self, cx, stage):
            self.cx, self.stage = cx, stage
            self.ms0, self.n0 = cx.stage_time_ms(stage)

        def done(self, bytes_per_call):
            ms1, n1 = self.cx.stage_time_ms(self.stage)
            calls = max(n1 - self.n0, 1)
            dev_ms = (ms1 - self.ms0) / calls
            ach = bytes_per_call / (dev_ms * 1e-3) / 1e9 if dev_ms > 0 else None
            return {"bound": "hbm", "algorithmic_bytes_per_call": int(bytes_per_call), "device_ms_per_call": dev_ms, "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": (ach / peak) if ach else None, "stage": lib.hso_stage_name(self.stage).decode()}

    # ---- F1: pyramid + stats, 64 frames per call -------------------------------------------------------------------------------
    nb = 64
    imgs = [pair["cur_img"]] * nb
    ptrs = (C.c_void_p * nb)(*[im.ctypes.data for im in imgs])
    ids = (C.c_int32 * nb)()
    integ = np.zeros(nb, np.float32)

    def up():
        ctx._chk(lib.hso_frame_upload_batch(ctx.h, nb, ptrs, W, H, W, ids, integ.ctypes.data_as(C.POINTER(C.c_float)), None))
        for i in range(nb):
            lib.hso_frame_release(ctx.h, ids[i])
    rf = Roof(ctx, 0)
    dt = timed(up, 5)
    roof_f1 = rf.done(nb * (W * H + (W * H) // 3))
    t0 = time.perf_counter()
    for _ in range(4):
        lv, _ = O.create_pyramid(pair["cur_img"], 5)
        for l in range(3):
            O.sobel5(lv[l])
        O.frame_stats(pair["cur_img"])
    cpu = (time.perf_counter() - t0) / 4
    out["frame_construction"] = {"gpu_frames_per_s_e2e": nb / dt, "cpu_frames_per_s_1core": 1.0 / cpu, "bytes_per_frame": W * H + (W * H) // 3, "roofline": roof_f1,
                                 "note": "hso_frame_upload_batch of 64 host images (H2D + k_pyr_tile + stats read-back) vs oracle pyramid+Sobel L0-L2+stats"}
    # ---- F3: 5000 candidates ------------------------------------------------------------------------------------------------------
    fid, _, _ = ctx.upload_frames([pair["ref_img"], pair["cur_img"]])
    M = 5000
    jobs = synth.make_align_jobs(args.seed + 6, pair, M=M, frac_edgelet=0.0)
    arr = (K.hso_align_job * M)()
    for m, j in enumerate(jobs):
        a = arr[m]
        a.ref_level, a.search_level, a.type, a.scale_patch = j["ref_level"], j["search_level"], j["type"], j["scale_patch"]
        for k in range(2):
            a.px_ref[k], a.grad[k], a.px_cur[k] = j["px_ref"][k], j["grad"][k], j["px_cur"][k]
        A = np.asarray(j["A_cur_ref"]).reshape(4)
        for k in range(4):
            a.A_cur_ref[k] = A[k]
        a.exposure_rat = j["exposure_rat"]
    refs = (C.c_int32 * M)(*([fid[0]] * M))
    res = (K.hso_align_result * M)()
    rf = Roof(ctx, 2)
    dt = timed(lambda: ctx._chk(lib.hso_align_batch(ctx.h, fid[1], M, arr, refs, 10, res)), 10)
    roof_f3 = rf.done(M * (580 + 24))
    rl, _ = O.create_pyramid(pair["ref_img"], 5)
    cl, _ = O.create_pyramid(pair["cur_img"], 5)
    sob = [O.sobel5(cl[l]) for l in range(3)]
    t0 = time.perf_counter()
    O.match_direct_batch(jobs[:2000], rl, cl, sob)
    cpu = (time.perf_counter() - t0) / 2000
    out["align_batch"] = {"gpu_patches_per_s_e2e": M / dt, "cpu_patches_per_s_1core": 1.0 / cpu, "M": M, "ms_per_call": dt * 1e3,
                          "algorithmic_bytes_per_patch": 580, "achieved_GBps_e2e": M * 580 / dt / 1e9, "roofline": roof_f3,
                          "note": "hso_align_batch (H2D jobs, k_align, D2H results) vs oracle match_direct_batch (python marshalling excluded from neither)"}
    # ---- F4: pose optimiser, 5000 features, 8 host frames; batch of 64 frames ----------------------------------------------------------
    probs = [synth.make_pose_problem(args.seed + 10 + i, args.cam, F=5000, K=8) for i in range(4)]
    batch = [probs[i % 4] for i in range(64)]
    pose_call, _ = ctx.pose_args(batch)  # flattened once; the timed call is hso_pose_optimize_batch itself on host buffers
    rf = Roof(ctx, 3)
    dt = timed(pose_call, 5)
    t0 = time.perf_counter()
    trials = sum(O.pose_optimize(p)["n_trials_total"] for p in probs)
    roof_f4 = rf.done(64 * 5000 * 72 * 2 * (trials / 4 + 2))  # 72 B per feature per pass, 2 passes per LM trial + the two scale / chi2 passes
    cpu = (time.perf_counter() - t0) / 4
    g = ctx.pose_optimize_batch(probs)
    out["pose_optimizer"] = {"gpu_frames_per_s_e2e": 64 / dt, "cpu_frames_per_s_1core": 1.0 / cpu, "features": 5000, "batch": 64,
                             "lm_trials_per_frame": trials / 4, "gpu_trials_per_frame": sum(r["n_trials_total"] for r in g) / 4, "roofline": roof_f4,
                             "note": "hso_pose_optimize_batch on flattened host arrays (H2D, k_pose_lm, D2H) vs oracle pose_optimize"}
    # ---- N2: FAST-9 detector on levels 0..2 of one frame (what fastDetectMT does per keyframe, feature_detection.cpp:498-514) ---------
    thr = 20
    fbuf = (K.hso_corner * 65536)()
    fcnt = C.c_int()

    fcnt3 = (C.c_int * 3)()

    def fast3():  # the C-ABI call itself with a caller-owned buffer (no Python list marshalling inside the timed loop): levels 0..2 in one call
        ctx._chk(lib.hso_fast_detect_levels(ctx.h, fid[1], 3, thr, 8, fbuf, 65536 // 3, fcnt3))
        return sum(fcnt3)
    n_c = fast3()
    rf = Roof(ctx, 6)
    dt = timed(fast3, 20)
    roof_n2 = rf.done(int(W * H * (1 + 1 / 4 + 1 / 16)) * 3 + n_c * 12)  # image read by score + score map written / read + corners out
    row = {"gpu_frames_per_s_e2e": 1.0 / dt, "gpu_ms_per_frame": dt * 1e3, "corners_after_nonmax": n_c, "levels": "0..2", "threshold": thr,
           "algorithmic_bytes_per_frame": int(W * H * (1 + 1 / 4 + 1 / 16)), "roofline": roof_n2, "note": "hso_fast_detect_levels: levels 0..2 in one call (nine kernels back to back, one synchronisation) incl. D2H of the corner lists"}
    if O.ref_fast_available():
        lv, _ = O.create_pyramid(pair["cur_img"], 5)
        t0 = time.perf_counter()
        for _ in range(5):
            for l in range(3):
                O.ref_fast9(lv[l], thr)
        cpu = (time.perf_counter() - t0) / 5
        row.update({"cpu_frames_per_s_1core": 1.0 / cpu, "cpu_kind": "reference (oracle/_ref/libfast_ref.so = thirdparty/fast compiled from the reference sources; "
                    "the reference itself runs the three levels on three threads)"})
    out["fast_detect"] = row
    # ---- N1: map reprojection + grid selection + findMatchDirect, 3000 map points against 4 keyframes --------------------------------------
    try:
        from hso_b200 import Context, make_cam
        sc = synth.make_reproject_scene(args.seed + 20, args.cam, M=3000, n_kf=4, max_fts=200)
        ctxs = Context(make_cam(W, H, c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), device=dev.index, max_frames=8, materialize_sobel=True)
        kf_ids, _, _ = ctxs.upload_frames(sc["kf_imgs"])
        cur_id = ctxs.upload_frames([sc["cur_img"]])[0][0]
        carr = Context.reproj_cands(sc["cands"], frame_ids=kf_ids)
        ctxs.reproject_match(cur_id, sc["T_cur_w"], sc["T_f_w"], carr, sc["grid"], sc["cell_order"])
        rf = Roof(ctxs, 4)
        t0 = time.perf_counter()
        for _ in range(20):
            _, summ = ctxs.reproject_match(cur_id, sc["T_cur_w"], sc["T_f_w"], carr, sc["grid"], sc["cell_order"])
        dt = (time.perf_counter() - t0) / 20
        roof_n1 = rf.done(3000 * (128 + 580 + 88))
        oc = (O.orc_reproj_cand * 3000).from_buffer_copy(bytes(Context.reproj_cands(sc["cands"])))
        pyrs = [O.create_pyramid(im, 5)[0] for im in sc["kf_imgs"]]
        cl2, _ = O.create_pyramid(sc["cur_img"], 5)
        sob2 = [O.sobel5(cl2[l]) for l in range(3)]
        og = O.orc_reproj_grid(*[sc["grid"][k] for k in ("cell_size", "n_cols", "n_rows", "max_fts", "align_max_iter")], 0)
        t0 = time.perf_counter()
        for _ in range(20):
            _, osum = O.reproject_match(sc["cam"], sc["T_cur_w"], sc["T_f_w"], oc, og, sc["cell_order"], 2, pyrs, cl2, sob2)
        cpu = (time.perf_counter() - t0) / 20
        t0 = time.perf_counter()
        O.reproject_speculative(sc["cam"], sc["T_cur_w"], sc["T_f_w"], oc, og, 2, pyrs, cl2, sob2)
        cpu_all = time.perf_counter() - t0
        out["reproject_match"] = {"gpu_frames_per_s_e2e": 1.0 / dt, "gpu_ms_per_frame": dt * 1e3, "cpu_frames_per_s_1core": 1.0 / cpu,
                                  "cpu_ms_per_frame": cpu * 1e3, "cpu_ms_all_candidates": cpu_all * 1e3, "candidates": 3000, "max_fts": 200,
                                  "gpu_matches": int(summ.n_matches), "gpu_trials": int(summ.n_trials), "cpu_trials": int(osum.n_trials), "roofline": roof_n1,
                                  "note": "hso_reproject_match (H2D candidates, k_reproject + k_align on ALL candidates + k_reproj_select, D2H) vs the "
                                          "oracle's sequential walk, which stops at the first match per cell and so aligns only ~n_trials candidates; "
                                          "cpu_ms_all_candidates = the oracle aligning every candidate like the device does"}
        ctxs.close()
    except Exception as e:  # a next-row diagnostic must not take the headline line down
        out["reproject_match"] = {"error": repr(e)}
    # ---- N4: raw 1280x1024 image -> ImageReader resize -> undistortion remap -> Frame (TUM monoVO wide, FOV model), 32 frames per call --------
    try:
        from hso_b200 import Context, make_cam
        cf = synth.CAMS["tum_fov"]
        ctxi = Context(make_cam(cf["width"], cf["height"], cf["fx"], cf["fy"], cf["cx"], cf["cy"], cf["d"], cf["model"]), device=dev.index, max_frames=40)
        rng = np.random.default_rng(args.seed + 30)
        raw = torch.from_numpy(np.stack([synth.texture(rng, 1280, 1024)] * 32)).pin_memory().numpy()
        raws = [raw[i] for i in range(32)]

        def up_raw():
            ids_, _, _ = ctxi.upload_raw_frames(raws, undistort=True)
            for i_ in ids_:
                ctxi.release(i_)
        up_raw()
        rf = Roof(ctxi, 0)
        t0 = time.perf_counter()
        for _ in range(5):
            up_raw()
        dt = (time.perf_counter() - t0) / 5
        roof_n4 = rf.done(32 * (1280 * 1024 + 2 * cf["width"] * cf["height"] * 4 + cf["width"] * cf["height"] * 4 // 3))
        t0 = time.perf_counter()
        m1, m2 = O.init_undistort_maps(cf)
        t_maps = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(3):
            im = O.remap_linear(O.resize_linear(raws[0], cf["width"], cf["height"]), m1, m2)
            lv, _ = O.create_pyramid(im, 5)
            O.frame_stats(im)
        cpu = (time.perf_counter() - t0) / 3
        out["raw_input"] = {"gpu_frames_per_s_e2e": 32 / dt, "gpu_ms_per_frame": dt * 1e3 / 32, "cpu_frames_per_s_1core": 1.0 / cpu, "raw": "1280x1024",
                            "frame": f"{cf['width']}x{cf['height']}", "bytes_per_frame": 1280 * 1024 + 2 * cf["width"] * cf["height"] * 4,
                            "cpu_map_build_ms": t_maps * 1e3, "roofline": roof_n4,
                            "note": "hso_frame_upload_raw_batch (H2D raw + k_resize_u8 + k_remap_u8 + pyramid + stats read-back) vs the oracle's "
                                    "cv::resize + cv::remap restatement + pyramid + stats (scalar C, not OpenCV's SIMD)"}
        ctxi.close()
    except Exception as e:
        out["raw_input"] = {"error": repr(e)}
    # ---- N3: depth-filter observation of 2000 seeds (3 keyframes) against one active frame ------------------------------------------------
    try:
        from hso_b200 import Context, make_cam
        sd = synth.make_depth_scene(args.seed + 40, args.cam, S=2000)
        ctxd = Context(make_cam(W, H, c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), device=dev.index, max_frames=8, materialize_sobel=True)
        kf_ids, _, _ = ctxd.upload_frames(sd["kf_imgs"])
        cur_id = ctxd.upload_frames([sd["cur_img"]])[0][0]
        sarr = Context.seed_obs(sd["seeds"], frame_ids=kf_ids)
        ctxd.depth_observe(cur_id, sd["T_cur_w"], sd["T_f_w"], sarr, sd["px_error_angle"])
        rf = Roof(ctxd, 5)
        t0 = time.perf_counter()
        for _ in range(20):
            gres = ctxd.depth_observe(cur_id, sd["T_cur_w"], sd["T_f_w"], sarr, sd["px_error_angle"])
        dt = (time.perf_counter() - t0) / 20
        roof_n3 = rf.done(2000 * (48 + 64 + 60 * 81))  # seed record in / out + ~60 scan steps and KLT iterations of 81 B of level image each
        oc = (O.orc_seed_obs * 2000).from_buffer_copy(bytes(Context.seed_obs(sd["seeds"])))
        pyrs = [O.create_pyramid(im, 5)[0] for im in sd["kf_imgs"]]
        cl3, _ = O.create_pyramid(sd["cur_img"], 5)
        sob3 = [O.sobel5(cl3[l]) for l in range(3)]
        t0 = time.perf_counter()
        for _ in range(5):
            O.depth_observe(sd["cam"], sd["T_cur_w"], sd["T_f_w"], oc, sd["px_error_angle"], pyrs, cl3, sob3)
        cpu = (time.perf_counter() - t0) / 5
        out["depth_observe"] = {"gpu_seeds_per_s_e2e": 2000 / dt, "gpu_ms_per_frame": dt * 1e3, "cpu_seeds_per_s_1core": 2000 / cpu,
                                "cpu_ms_per_frame": cpu * 1e3, "seeds": 2000, "updated": int(sum(gres[i].res == 1 for i in range(2000))), "roofline": roof_n3,
                                "note": "hso_depth_observe (H2D seeds, k_depth_observe, D2H results) vs the oracle's observeDepthRow on one core "
                                        "(the reference splits the seed list over 4 threads)"}
        ctxd.close()
    except Exception as e:
        out["depth_observe"] = {"error": repr(e)}
    # ---- a13b: seed stage of reprojectMap, 2000 seeds -----------------------------------------------------------------------------------------
    try:
        from hso_b200 import Context, make_cam
        ss = synth.make_seed_reproject_scene(args.seed + 50, args.cam, S=2000, max_fts=200)
        ctxs2 = Context(make_cam(W, H, c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), device=dev.index, max_frames=8, materialize_sobel=True)
        kf_ids, _, _ = ctxs2.upload_frames(ss["kf_imgs"])
        cur_id = ctxs2.upload_frames([ss["cur_img"]])[0][0]
        sarr = Context.seed_obs(ss["seeds"], frame_ids=kf_ids)
        ctxs2.reproject_seeds(cur_id, ss["T_cur_w"], ss["T_f_w"], sarr, ss["grid"], ss["cell_order"], n_matches_in=40)
        rf = Roof(ctxs2, 4)
        t0 = time.perf_counter()
        for _ in range(20):
            _, ssum = ctxs2.reproject_seeds(cur_id, ss["T_cur_w"], ss["T_f_w"], sarr, ss["grid"], ss["cell_order"], n_matches_in=40)
        dt = (time.perf_counter() - t0) / 20
        roof_s = rf.done(2000 * (96 + 580 + 88))
        oc = (O.orc_seed_obs * 2000).from_buffer_copy(bytes(Context.seed_obs(ss["seeds"])))
        pyrs = [O.create_pyramid(im, 5)[0] for im in ss["kf_imgs"]]
        cl4, _ = O.create_pyramid(ss["cur_img"], 5)
        sob4 = [O.sobel5(cl4[l]) for l in range(3)]
        og = O.orc_reproj_grid(*[ss["grid"][k] for k in ("cell_size", "n_cols", "n_rows", "max_fts", "align_max_iter")], 0)
        t0 = time.perf_counter()
        for _ in range(10):
            _, osum = O.reproject_seeds(ss["cam"], ss["T_cur_w"], ss["T_f_w"], oc, og, ss["cell_order"], 40, 2, pyrs, cl4, sob4)
        cpu = (time.perf_counter() - t0) / 10
        out["reproject_seeds"] = {"gpu_ms_per_frame": dt * 1e3, "cpu_ms_per_frame": cpu * 1e3, "seeds": 2000, "gpu_matches": int(ssum.n_matches),
                                  "gpu_trials": int(ssum.n_trials), "cpu_trials": int(osum.n_trials), "roofline": roof_s,
                                  "note": "hso_reproject_seeds (k_reproject_seed + k_align on ALL seeds + k_seed_select) vs the oracle's sequential walk"}
        ctxs2.close()
    except Exception as e:
        out["reproject_seeds"] = {"error": repr(e)}
    for f_ in fid:
        ctx.release(f_)
    return out


def single_stream(args, device, torch):
    """The reference's real operating point: ONE stream, one frame at a time through the synchronous C-ABI — at ~200 features per frame
    (Config::maxFts) and at the metric's 3000 patches — beside the single-threaded CPU implementation on the same frames. Per mode and patch
    count: ms per frame through hso_add_frames_track_batch(B = 1) with argument records built once (as a C++ caller holding long-lived Frame /
    Feature objects would), LM trials per frame, microseconds per LM trial (the latency floor of a serial trial sequence), the CPU's ms per frame
    on one core and the ratio."""
    from hso_b200 import Context, make_cam, _capi as K
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    out = {}
    for F in (200, args.patches):
        probs = build_workload(6, F, args.cam, args.seed + 99, 0)
        c = probs[0]["cam"]
        ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), device=device, max_frames=8,
                      max_features=max(8192, F))
        row = {}
        for mode, ic in (("forward", False), ("inverse-compositional", True)):
            for p in probs:
                if "_rid" not in p:
                    ids, integ, _ = ctx.upload_frames([p["ref_img"]])
                    p["_rid"], p["_rint"] = ids[0], integ[0]
            prm = K.hso_track_params(int(ic), 4, args.min_level, args.n_iter)
            recs = []
            for p in probs:
                jarr, keep = ctx._track_jobs([dict(ref=p["_rid"], cur=0, px=p["px"], f=p["f"], dist=p["dist"], T_cur_ref=p["T0"], exposure_rat=-1.0)])
                img = np.ascontiguousarray(p["cur_img"])
                recs.append((jarr, keep, img, (C.c_void_p * 1)(img.ctypes.data)))
            nid, res1 = (C.c_int32 * 1)(), (K.hso_track_result * 1)()
            Wc, Hc = c["width"], c["height"]

            def frame1(r):
                ctx._chk(ctx.lib.hso_add_frames_track_batch(ctx.h, C.byref(prm), 1, r[3], Wc, Hc, Wc, r[0], nid, None, None, res1))
                ctx.lib.hso_frame_release(ctx.h, nid[0])
                return res1[0].n_iters
            for r in recs:
                frame1(r)
            reps = 20 if F <= 500 else 8
            t0 = time.perf_counter()
            its = 0
            for _ in range(reps):
                for r in recs:
                    its += frame1(r)
            dt = time.perf_counter() - t0
            n_fr = reps * len(recs)
            it_c, dt_c, kind = run_cpu(probs, ic, 1)
            row[mode] = {"gpu_ms_per_frame": 1e3 * dt / n_fr, "gpu_frames_per_s": n_fr / dt, "gpu_iterations_per_s": its / dt,
                         "lm_trials_per_frame": its / n_fr, "gpu_us_per_lm_trial": 1e6 * dt / max(its, 1),
                         "cpu_ms_per_frame_1core": 1e3 * dt_c / len(probs), "cpu_kind": kind, "speedup_vs_1core": (dt_c / len(probs)) / (dt / n_fr)}
        for p in probs:
            p.pop("_rid", None)
        ctx.close()
        out[f"F={F}"] = row
    return out

# ---------------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2368, help="independent frame pairs per GPU per step (16 per SM: a launch ends with SMs idling behind "
                    "the last problems, and more, hence relatively shorter, units shrink that tail; measured 3.89 / 4.06 / 4.18 M it/s at 1184 / 1776 / 2368)")
    ap.add_argument("--patches", type=int, default=3000)
    ap.add_argument("--cam", default="icl", choices=list(synth.CAMS))
    ap.add_argument("--ic", action="store_true", help="inverse-compositional mode")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--wide-features", action="store_true", help="e2e input as px / f / dist (48 B per feature) instead of the compact xyz / px32 layout (32 B)")
    ap.add_argument("--e2e-debug", action="store_true", help="time the stages of the pipelined e2e call alone and print one traced call (stderr)")
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--shape", default="", help="per-level launch shape overrides 'level:ctas:threads,...' (tuning)")
    ap.add_argument("--seed", type=int, default=0x450)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-rows", action="store_true")
    ap.add_argument("--no-ic-dual", action="store_true")
    ap.add_argument("--min-level", type=int, default=1, help="lowest pyramid level (0 = the reference's relocalisation setting, with --n-iter 15)")
    ap.add_argument("--n-iter", type=int, default=50)
    ap.add_argument("--pipe", default="", help="chunk:streams of the pipelined e2e call (hso_set_pipeline; tuning)")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--pageable-features", action="store_true", help="keep the feature arrays in pageable memory (host-side flattening path)")
    ap.add_argument("--parity-samples", type=int, default=16, help="problems of the timed batch whose final pose is checked against the oracle")
    args = ap.parse_args()
    claim_stdout()
    TRACK_CFG.update(min_level=args.min_level, n_iter=args.n_iter)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    import torch
    import torch.distributed as dist
    from hso_b200 import Context, make_cam, _capi as K

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, F = args.batch, args.patches
    probs = build_workload(B, F, args.cam, args.seed, rank)
    c = probs[0]["cam"]
    W, H = c["width"], c["height"]
    ctx = Context(make_cam(W, H, c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), device=local_rank, max_frames=2 * B + 72,
                  max_features=max(8192, F))
    lib = ctx.lib
    if args.cluster or args.threads:
        ctx.set_cluster(args.cluster, args.threads)
    if args.no_ic_dual:
        ctx._chk(lib.hso_track_set_ic_dual(ctx.h, 0))
    if args.pipe:
        ch, st = (int(v) for v in args.pipe.split(":"))
        ctx._chk(lib.hso_set_pipeline(ctx.h, ch, st))
    for item in [x for x in args.shape.split(",") if x]:
        lv, cc, th = (int(v) for v in item.split(":"))
        ctx._chk(lib.hso_track_set_level_shape(ctx.h, lv, cc, th))

    # pinned host copies of every image (e2e uploads from these), device-resident raw current images (for `value`)
    host_cur = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
    host_ref = torch.empty((len(set(p["base"] for p in probs)), H, W), dtype=torch.uint8).pin_memory()
    for b, p in enumerate(probs):
        host_cur[b].copy_(torch.from_numpy(p["cur_img"]))
        host_ref[p["base"]].copy_(torch.from_numpy(p["ref_img"]))
    ref_np = [host_ref[p["base"]].numpy() for p in probs]
    cur_np = [host_cur[b].numpy() for b in range(B)]
    ref_ids, ref_int, _ = ctx.upload_frames(ref_np)
    cur_ids, cur_int, _ = ctx.upload_frames(cur_np)
    dev_cur = host_cur.to(dev, non_blocking=False)  # [B,H,W] u8, rows 16-byte aligned (W % 16 == 0 for the half-sample path)
    dev_ptrs = (C.c_void_p * B)(*[dev_cur[b].data_ptr() for b in range(B)])
    cur_ids_c = (C.c_int32 * B)(*cur_ids)

    # the caller's feature arrays: one pinned blob per field for the whole batch ([B][F][2], [B][F][3], [B][F] doubles), the layout a batched
    # caller keeps; pinned, they are DMA-copied as they are and flattened on the device (hso_track_set_direct_inputs, default auto)
    if args.pageable_features:
        px_all, f_all, dist_all = np.empty((B, F, 2)), np.empty((B, F, 3)), np.empty((B, F))
    else:
        px_all = torch.empty((B, F, 2), dtype=torch.float64).pin_memory().numpy()
        f_all = torch.empty((B, F, 3), dtype=torch.float64).pin_memory().numpy()
        dist_all = torch.empty((B, F), dtype=torch.float64).pin_memory().numpy()
    jobs = []
    for b, p in enumerate(probs):
        px_all[b], f_all[b], dist_all[b] = p["px"], p["f"], p["dist"]
        a0 = float(np.float32(cur_int[b]) / np.float32(ref_int[b]))
        jobs.append(dict(ref=ref_ids[b], cur=cur_ids[b], px=px_all[b], f=f_all[b], dist=dist_all[b], T_cur_ref=p["T0"], exposure_rat=a0))
    # The e2e call's default input is the compact layout of hso_track_job (32 B per feature instead of 48; the call is bound by the host->device
    # copies): what a caller's gather loop over ref_frame->fts_ writes — xyz = f * dist (src/CoarseTracker.cpp:292), px as float32 (exact for the
    # tracker), features without a point left out. One pinned blob per field for the whole batch. Results are bit-identical to the wide layout
    # (tests/test_gpu_track.py::test_compact_feature_layout_is_bit_identical); --wide-features times the 48-byte layout.
    jobs_e2e = jobs
    if not args.wide_features and not args.pageable_features:
        nval = [int((p["dist"] >= 0).sum()) for p in probs]
        offs = np.concatenate([[0], np.cumsum(nval)])
        xyz_all = torch.empty((int(offs[-1]), 3), dtype=torch.float64).pin_memory().numpy()
        px32_all = torch.empty((int(offs[-1]), 2), dtype=torch.float32).pin_memory().numpy()
        jobs_e2e = []
        for b, p in enumerate(probs):
            xyz, px32 = Context.compact_features(p["px"], p["f"], p["dist"])
            xyz_all[offs[b]:offs[b + 1]], px32_all[offs[b]:offs[b + 1]] = xyz, px32
            jobs_e2e.append(dict(ref=ref_ids[b], cur=cur_ids[b], xyz=xyz_all[offs[b]:offs[b + 1]], px32=px32_all[offs[b]:offs[b + 1]], T_cur_ref=p["T0"],
                                 exposure_rat=jobs[b]["exposure_rat"]))

    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    levels = list(range(4, args.min_level - 1, -1))

    def device_step():
        ctx._chk(lib.hso_frame_rebuild_batch_device(ctx.h, B, dev_ptrs, W, H, W, cur_ids_c))
        ctx.track_run()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM -------------------------------------------------------------------------------------
    ctx.track_stage(jobs, inverse_comp=args.ic, max_level=4, min_level=args.min_level, n_iter=args.n_iter)
    sampler = ClockSampler(local_rank)
    sampler.start()  # before the warm-up: nvidia-smi's first line takes ~0.1 s, the timed region may be shorter
    for _ in range(args.warmup):
        device_step()
    ctx.synchronize()
    out = ctx.track_collect()
    iters_per_step = sum(out[b].n_iters for b in range(B))
    patch_evals = {l: sum(out[b].visible_patch_evals[l] for b in range(B)) for l in levels}
    cyc = [sum(out[b].cycles[k] for b in range(B)) for k in range(8)]
    diag = {"setup_frac_of_kernel_cycles": cyc[1] / max(cyc[0], 1), "refpatch_frac_of_kernel_cycles": cyc[3] / max(cyc[0], 1), "threshold_residuals_frac": cyc[4] / max(cyc[0], 1),
            "median_select_frac": cyc[5] / max(cyc[0], 1), "pivoted_solves": cyc[7] & 0xFFFFFFFF, "select_fallbacks_per_problem_level": (cyc[7] >> 32) / (len(levels) * B), "control_cycles_per_iteration": cyc[2] / max(iters_per_step + len(levels) * B, 1),
            "kernel_cycles_per_problem_level": cyc[0] / (len(levels) * B), "mad_select_frac": cyc[6] / max(cyc[0], 1), "serial_control_frac_of_kernel_cycles": cyc[2] / max(cyc[0], 1),
            "iters_per_problem": {"mean": iters_per_step / B, "max": max(out[b].n_iters for b in range(B)), "min": min(out[b].n_iters for b in range(B))}}
    ctx._chk(lib.hso_track_set_profile(ctx.h, 1))
    launches0 = ctx.kernel_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        device_step()
    e1.record(stream)
    barrier()
    t_region1 = time.perf_counter()
    clocks = sampler.stop(t_region0, t_region1)
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - launches0
    lvl_ms = {}
    for l in levels:
        ms, n = C.c_double(), C.c_uint64()
        ctx._chk(lib.hso_track_level_profile(ctx.h, l, C.byref(ms), C.byref(n)))
        lvl_ms[l] = ms.value / max(n.value, 1)
    ctx._chk(lib.hso_track_set_profile(ctx.h, 0))
    final_dev = ctx.track_collect()  # results of the last timed device-resident step (checked against the oracle below)

    # ---- e2e: host buffers through the C-ABI ------------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # the timed calls are the C-ABI entry points themselves on HOST buffers; the argument records are built once (the reference
        # caller owns long-lived Frame/Feature objects too) and only the per-step fields are refreshed
        prm = K.hso_track_params(int(args.ic), 4, args.min_level, args.n_iter)
        jarr, keep = ctx._track_jobs(jobs_e2e)
        img_ptrs = (C.c_void_p * B)(*[im.ctypes.data for im in cur_np])
        new_ids = (C.c_int32 * B)()
        integ = np.zeros(B, np.float32)
        res = (K.hso_track_result * B)()
        ref_int32 = np.asarray(ref_int, np.float32)
        fptr = C.POINTER(C.c_float)

        gmean = np.zeros(B, np.float32)
        for b in range(B):
            jarr[b].exposure_rat = -1.0  # the device forms cur.integralImage_/ref.integralImage_ (CoarseTracker.cpp:60)

        ids_live, ids_next = cur_ids_c, new_ids

        def e2e_step():
            nonlocal ids_live, ids_next
            lib.hso_frame_release_batch(ctx.h, B, ids_live)  # the previous step's frames (~Frame)
            # ONE public call on host buffers: H2D of B images and of the flattened feature arrays, pyramids + statistics, CoarseTracker
            # L4->L1, D2H of the results and statistics (chunk-pipelined inside: copies of chunk c+1 overlap the kernels of chunk c)
            ctx._chk(lib.hso_add_frames_track_batch(ctx.h, C.byref(prm), B, img_ptrs, W, H, W, jarr, ids_next, integ.ctypes.data_as(fptr),
                                                    gmean.ctypes.data_as(fptr), res))
            ids_live, ids_next = ids_next, ids_live
            return 0
        for _ in range(2):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        n_e2e_steps = max(2, args.steps)
        it_e2e = 0
        for _ in range(n_e2e_steps):
            e2e_step()
        f1.record(stream)
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        if args.e2e_debug and rank == 0:  # tuning aid: the stages of the pipelined call alone (HSO_PIPE_DEBUG) and one traced call (HSO_PIPE_TRACE)
            for name, flag in (("copies only", "1"), ("kernels only", "2"), ("everything", "0")):
                os.environ["HSO_PIPE_DEBUG"] = flag
                tt = []
                for _ in range(6):
                    torch.cuda.synchronize(); t0 = time.perf_counter(); e2e_step(); torch.cuda.synchronize(); tt.append(1e3 * (time.perf_counter() - t0))
                print(f"[e2e-debug] {name}: median {np.median(tt[2:]):.2f} ms min {min(tt):.2f} ms", file=sys.stderr)
            os.environ["HSO_PIPE_DEBUG"] = "0"
            os.environ["HSO_PIPE_TRACE"] = "1"
            e2e_step()
            del os.environ["HSO_PIPE_TRACE"]
        it_e2e = n_e2e_steps * sum(res[b].n_iters for b in range(B))  # every step runs the same problems: counted once, outside the timed region
        if ids_live is not cur_ids_c:
            C.memmove(cur_ids_c, ids_live, C.sizeof(cur_ids_c))  # later sections rebuild the frames these ids name
        nvalid = sum(int((p["dist"] >= 0).sum()) for p in probs)
        if args.pageable_features:  # flattened on the host: 5 doubles per feature with a depth (padded to 32 features)
            h2d = B * W * H + sum(40 * max(32, (int((p["dist"] >= 0).sum()) + 31) // 32 * 32) for p in probs) + B * (96 + 4 + 96)
        elif args.wide_features:    # direct: the caller's 6 doubles per feature as they are
            h2d = B * W * H + 48 * B * F + B * (96 + 4 + 128)
        else:                       # compact: 3 doubles + 2 floats per feature with a point
            h2d = B * W * H + 32 * nvalid + B * (96 + 4 + 128)
        d2h = B * (C.sizeof(K.hso_track_result) + 8)  # results + {integralImage_, gradMean_}
        e2e = dict(ms=ms_e2e, steps=n_e2e_steps, iters=it_e2e, h2d=h2d, d2h=d2h)

    # ---- parity of the timed batch: final poses of sampled problems against the CPU oracle (checker only, outside every timed region) ----------
    parity = None
    if rank == 0 and not args.no_parity_check:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        O.load()
        final = final_dev
        n_chk = min(B, args.parity_samples)
        sample = [int(round(i * (B - 1) / max(n_chk - 1, 1))) for i in range(n_chk)]
        pyr = {}
        worst = worst_a = worst_e2e = 0.0
        for b in sample:
            p = probs[b]
            if p["base"] not in pyr:
                pyr[p["base"]] = (O.create_pyramid(p["ref_img"], 5)[0], O.create_pyramid(p["cur_img"], 5)[0])
            tpo = O.TrackProblem(c, pyr[p["base"]][0], pyr[p["base"]][1], p["px"], p["f"], p["dist"])
            ro = tpo.run(p["T0"], jobs[b]["exposure_rat"], inverse_comp=args.ic)
            Tg = np.array(final[b].T_cur_ref[:]).reshape(3, 4)
            worst = max(worst, float(np.abs(Tg - ro["T_cur_ref"]).max()))
            worst_a = max(worst_a, abs(float(final[b].exposure_rat) - ro["exposure_rat"]))
            if e2e:
                worst_e2e = max(worst_e2e, float(np.abs(np.array(res[b].T_cur_ref[:]).reshape(3, 4) - ro["T_cur_ref"]).max()))
        tol = 2e-4
        parity = {"n": n_chk, "tol_abs_pose_entry": tol, "max_abs_pose_diff": worst, "max_abs_exposure_diff": worst_a,
                  "e2e_max_abs_pose_diff": worst_e2e if e2e else None, "ok": bool(worst < tol and worst_a < tol and worst_e2e < tol),
                  "against": "oracle/liboracle_hso.so CoarseTracker::run restatement, same inputs (final pose of the timed batch's problems)"}
        if not parity["ok"]:
            print(f"bench.py: PARITY FAILURE against the oracle: {parity}", file=sys.stderr)

    # ---- reduce over ranks: max time, summed work -----------------------------------------------------------------------------------
    stat = torch.tensor([ms_total, float(iters_per_step * args.steps), float(B * args.steps), e2e["ms"] if e2e else 0.0,
                         float(e2e["iters"]) if e2e else 0.0, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, iters_all, frames_all = mx[0].item(), sm[1].item(), sm[2].item()
        ms_e2e_all, it_e2e_all, launches_all = mx[3].item(), sm[4].item(), sm[5].item()
    else:
        iters_all, frames_all = stat[1].item(), stat[2].item()
        ms_e2e_all, it_e2e_all, launches_all = stat[3].item(), stat[4].item(), stat[5].item()

    if rank == 0:
        peak, peak_src = measured_peak()
        tab = BYTES_PER_PATCH_EVAL_IC if args.ic else BYTES_PER_PATCH_EVAL
        dom = max(levels, key=lambda l: lvl_ms[l])
        alg_bytes = {l: patch_evals[l] * tab[l] for l in levels}
        achieved = alg_bytes[dom] / (lvl_ms[dom] * 1e-3) / 1e9
        all_ach = sum(alg_bytes.values()) / (sum(lvl_ms.values()) * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "track_l1_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))  # ncu --set full capture of the level-1 kernel (tools/ncu_summary.py); DRAM bytes scale with the batch
                traffic = tj["dram_bytes_per_problem"] * B if "dram_bytes_per_problem" in tj else tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "CoarseTracker LM iterations/sec @640x480, 3k patches", "value": iters_all / (ms_total * 1e-3), "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "frames_per_s": frames_all / (ms_total * 1e-3),
            "config": {"workload": f"{args.cam} {W}x{H}, {F} patches/frame, {B} independent frame pairs per GPU per step: pyramid+stats of the "
                                   f"current image then CoarseTracker L4->L{args.min_level} n_iter={args.n_iter} {'inverse-compositional' if args.ic else 'forward'}, natural convergence",
                       "batch_per_gpu": B, "patches": F, "lm_iterations_per_step_per_gpu": iters_per_step,
                       "l2": f"inputs larger than L2: {B} x (2 pyramids + feature scratch) = {B * (2 * 410000 + F * 25 * 8 + F * 40) / 1e6:.0f} MB per step vs 126 MB L2",
                       "parallelism": f"{world} independent batch(es), one per GPU, no data-path collective", "numa_binding_rank0": numa},
            "gpu_launches": int(launches_all), "diag": diag, "parity_checked": parity["n"] if parity else 0, "parity": parity,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": f"k_track_level (level {dom})", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_bytes[dom], "kernel_ms": lvl_ms[dom],
                         "share_of_step": lvl_ms[dom] / (ms_total / args.steps),
                         "all_levels": {"achieved": all_ach, "frac": all_ach / peak, "kernel_ms": {str(l): lvl_ms[l] for l in levels},
                                        "algorithmic_bytes": {str(l): alg_bytes[l] for l in levels}},
                         "note": ("algorithmic bytes = visible patch evaluations x SURVEY 8(d) bytes/patch/eval; the cached intensity + gradient planes of "
                                  "the resident problems exceed the L2 and are streamed from HBM every evaluation, so DRAM traffic is of the order of the "
                                  "algorithmic figure (HBM and issue bound)") if args.ic else
                                 ("algorithmic bytes = visible patch evaluations x SURVEY 8(d) bytes/patch/eval; the level image is staged in shared "
                                  "memory, the streamed reference-patch cache and the scratch stay in L2, so DRAM traffic is far below the algorithmic "
                                  "figure (issue bound, not HBM bound)")},
        }
        if e2e:
            line["e2e"] = {"value": it_e2e_all / (ms_e2e_all * 1e-3), "unit": "iterations/s", "h2d_bytes_per_step": e2e["h2d"] * world,
                           "d2h_bytes_per_step": e2e["d2h"] * world, "frames_per_s": world * B * e2e["steps"] / (ms_e2e_all * 1e-3),
                           "ms_per_step": ms_e2e_all / e2e["steps"], "steps": e2e["steps"],
                           "features": ("px/f/dist doubles, 48 B per feature, pageable (flattened by the host pool)" if args.pageable_features else
                                        "px/f/dist doubles, 48 B per feature, pinned (copied as they are, flattened on the device)" if args.wide_features else
                                        "compact hso_track_job layout, 32 B per feature with a point: xyz = f*dist doubles + px float32, pinned")}
        if world == 1 and not args.no_cpu_baseline:
            cores = 1
            sample = build_workload(256, F, args.cam, args.seed, 0)
            run_cpu(sample[:1], args.ic, 1)
            it, dt, kind = run_cpu(sample, args.ic, 1)
            line["cpu_baseline"] = {"value": it / dt, "unit": "iterations/s", "cores": cores, "kind": kind,
                                    "sample": f"256 frames of the same workload ({it} LM iterations, {dt:.1f} s) single-threaded like the reference's "
                                              f"tracking thread, " + ("the reference's own sources compiled unmodified (oracle/_ref/libhso_ref.so)"
                                                                      if kind == "reference" else "oracle port") + f"; host has {os.cpu_count()} cores",
                                    "frames_per_s": 256 / dt}
            it_p, dt_p, _ = run_cpu(sample[:64], args.ic, 1, kind="port")
            line["cpu_baseline"]["port_value"] = it_p / dt_p  # the oracle restatement on the same host core, for comparison
        if world == 1 and not args.no_other_rows:
            # the same batch in the other Jacobian mode (the reference picks inverse-compositional unless the new frame's gradients
            # got stronger, src/frame_handler_mono.cpp:184-203), device-resident like `value`
            ctx.track_stage(jobs, inverse_comp=not args.ic, max_level=4, min_level=args.min_level, n_iter=args.n_iter)
            for _ in range(2):
                device_step()
            ctx.synchronize()
            ctx._chk(lib.hso_track_set_profile(ctx.h, 1))
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            for _ in range(args.steps):
                device_step()
            g1.record(stream)
            ctx.synchronize()
            lvl2 = {}
            for l in levels:
                ms, n = C.c_double(), C.c_uint64()
                ctx._chk(lib.hso_track_level_profile(ctx.h, l, C.byref(ms), C.byref(n)))
                lvl2[l] = ms.value / max(n.value, 1)
            ctx._chk(lib.hso_track_set_profile(ctx.h, 0))
            o2 = ctx.track_collect()
            it2 = sum(o2[b].n_iters for b in range(B))
            tab2 = BYTES_PER_PATCH_EVAL if args.ic else BYTES_PER_PATCH_EVAL_IC
            alg2 = {l: sum(o2[b].visible_patch_evals[l] for b in range(B)) * tab2[l] for l in levels}
            dom2 = max(levels, key=lambda l: lvl2[l])
            peak2, _ = measured_peak()
            line["other_mode"] = {"mode": "forward" if args.ic else "inverse-compositional", "value": it2 * args.steps / (g0.elapsed_time(g1) * 1e-3),
                                  "unit": "iterations/s", "ms_per_step": g0.elapsed_time(g1) / args.steps,
                                  "roofline": {"bound": "hbm", "kernel": f"k_track_level (level {dom2})", "unit": "GB/s", "peak": peak2,
                                               "achieved": alg2[dom2] / (lvl2[dom2] * 1e-3) / 1e9, "frac": alg2[dom2] / (lvl2[dom2] * 1e-3) / 1e9 / peak2,
                                               "all_levels_frac": sum(alg2.values()) / (sum(lvl2.values()) * 1e-3) / 1e9 / peak2,
                                               "kernel_ms": {str(l): lvl2[l] for l in levels}}}
            line["other_rows"] = other_rows(ctx, lib, args, dev, torch, K)
            line["single_stream"] = single_stream(args, local_rank, torch)
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
