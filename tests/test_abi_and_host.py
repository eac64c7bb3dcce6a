"""CPU tests: the C-ABI library loads and exports every symbol include/hso_b200.h declares; struct layouts agree between the
header, the ctypes binding and the oracle's records; host-side logic (synthetic generator, bench accounting, multi-rank reduce
over gloo). No compute call needs a GPU here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hso_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hso_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hso_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH), "libhso_b200.so missing: run __graft_entry__.build()"
    lib = C.CDLL(_capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hso_b200.h but not exported"
    # and the python binding covers exactly the header
    assert sorted(_capi.SYMBOLS) == declared
    _capi.load()


def test_no_device_is_reported_not_faked():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hso_b200 import Context, HsoError, make_cam
    with pytest.raises(HsoError) as e:
        Context(make_cam(640, 480, 480, 480, 320, 240))
    assert e.value.code == -3  # HSO_ERR_NO_DEVICE: there is no CPU fallback


def test_struct_sizes_match_header_layout():
    from hso_b200 import _capi as K
    import oracle_lib as O
    # sizes computed from the header's field lists (natural alignment)
    assert C.sizeof(K.hso_cam) == 16 + 4 * 8 + 5 * 8
    assert C.sizeof(K.hso_cfg) == 32
    assert C.sizeof(K.hso_track_job) == 16 + 3 * 8 + 96 + 8 + 2 * 8  # + the compact layout's two pointers (xyz, px32)
    assert C.sizeof(K.hso_trace) == 8 + 96 + 8 + 8 * (49 + 7 + 7 + 1) + 12 + 8 + 4
    assert C.sizeof(K.hso_align_job) == C.sizeof(O.orc_align_job) == 16 + 8 * 10 + 8
    assert C.sizeof(K.hso_align_result) == C.sizeof(O.orc_align_result) == 32
    assert C.sizeof(K.hso_pose_result) == C.sizeof(O.orc_pose_result)
    assert C.sizeof(K.hso_trace) == C.sizeof(O.orc_trace)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hso_b200")
    for dp_, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp_, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "hso_oracle.h" not in txt, f"{f} references the oracle"


def test_synth_is_seeded_and_consistent():
    from hso_b200 import synth
    a, b = synth.make_pair(5, "icl", F=50), synth.make_pair(5, "icl", F=50)
    assert np.array_equal(a["ref_img"], b["ref_img"]) and np.array_equal(a["px"], b["px"])
    # ray * dist lies on the plane z = depth
    ok = a["dist"] > 0
    assert np.allclose((a["f"] * a["dist"][:, None])[ok, 2], 4.0)
    assert a["ref_img"].dtype == np.uint8 and a["ref_img"].shape == (480, 640)


def test_bench_reference_arm_and_gloo_reduce(tmp_path):
    # the reference arm is pure CPU: run it on a tiny workload and check the JSON contract
    env = dict(os.environ)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--patches", "200"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "iterations/s" and line["higher_is_better"] is True
    import ref_lib
    # the reference's own compiled sources when oracle/_ref/libhso_ref.so is present (built where /root/reference exists), else the oracle port
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_lib.available() else "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    # multi-rank bookkeeping (max over ranks of the time, sum over ranks of the work) over gloo, world_size 2
    script = tmp_path / "reduce.py"
    script.write_text(
        "import os, torch, torch.distributed as dist\n"
        "dist.init_process_group('gloo')\n"
        "r = dist.get_rank()\n"
        "stat = torch.tensor([10.0 + 5 * r, 100.0 * (r + 1), 8.0], dtype=torch.float64)\n"
        "mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)\n"
        "sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)\n"
        "dist.barrier()\n"
        "if r == 0: print('OK', mx[0].item(), sm[1].item(), sm[2].item())\n"
        "dist.destroy_process_group()\n")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29671", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK 15.0 300.0 16.0" in out.stdout
    # rank != 0 of the reference arm exits 0 without work
    env2 = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=60, env=env2)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_host_layer_wire_formats(tmp_path):
    """ImageReader's four timestamp line formats and the TUM trajectory line of saveResult (row N4), C++ host layer, CPU only."""
    import subprocess
    from hso_b200 import _capi
    exe = tmp_path / "host_formats"
    libdir = os.path.dirname(_capi.LIB_PATH)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(root, "tests", "cpp", "host_formats.cpp"), "-o", str(exe), f"-L{libdir}", "-lhso_b200",
                           f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stderr


def test_host_make_depth_ref_matches_oracle(tmp_path):
    """Row a4: hso::b200::CoarseTracker::makeDepthRef (the pointer-chasing host half of src/CoarseTracker.cpp:210-240) against the oracle's
    orc_make_depth_ref and an independent numpy statement: 5 host keyframes with distinct non-identity poses, a non-identity reference pose,
    features without a point, points behind the reference camera and points with z below the 1e-5 cut."""
    import subprocess
    import oracle_lib as O
    from hso_b200 import _capi, synth
    rng = np.random.default_rng(404)
    F, K = 400, 5
    T_ref = synth.se3_exp(np.array([0.3, -0.2, 0.1, 0.05, -0.08, 0.12]))
    T_hosts = [synth.se3_exp(np.concatenate([rng.normal(0, 0.4, 3), rng.normal(0, 0.15, 3)])) for _ in range(K)]
    host_of = rng.integers(0, K, F).astype(np.int32)
    has = (rng.uniform(size=F) < 0.85).astype(np.int32)
    ray = np.stack([rng.uniform(-0.6, 0.6, F), rng.uniform(-0.5, 0.5, F), np.ones(F)], axis=1)
    f_host = ray / np.linalg.norm(ray, axis=1, keepdims=True)
    idist = 1.0 / rng.uniform(0.5, 8.0, F)
    # force the rejections: points whose position in the reference frame has z < 1e-5 (behind the camera, and just below the threshold)
    T_r_h = [T_ref @ np.linalg.inv(T) for T in T_hosts]
    forced = []
    for i in range(0, 60, 3):
        T = np.linalg.inv(T_r_h[host_of[i]])
        z = -1.0 if i % 2 else 0.5e-5
        p_ref = np.array([0.1, -0.05, z])
        p_h = T[:3, :3] @ p_ref + T[:3, 3]
        d = np.linalg.norm(p_h)
        f_host[i], idist[i], has[i] = p_h / d, 1.0 / d, 1
        forced.append(i)
    blob = tmp_path / "depthref.bin"
    with open(blob, "wb") as fh:
        fh.write(np.array([F, K], np.int32).tobytes())
        fh.write(np.ascontiguousarray(T_ref[:3], np.float64).tobytes())
        fh.write(np.ascontiguousarray(np.stack([T[:3] for T in T_hosts]), np.float64).tobytes())
        fh.write(np.ascontiguousarray(f_host, np.float64).tobytes()); fh.write(idist.astype(np.float64).tobytes())
        fh.write(has.tobytes()); fh.write(host_of.tobytes())
    exe = tmp_path / "host_depthref"
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "host_depthref.cpp"), "-o", str(exe), f"-L{libdir}", "-lhso_b200",
                           f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe), str(blob)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    got = np.array([float(x) for x in out.stdout.split()])
    exp = O.make_depth_ref(T_ref[:3], has, f_host, idist, np.stack([T_hosts[h][:3] for h in host_of]))
    # independent numpy statement
    ref = -np.ones(F)
    for i in range(F):
        if not has[i]:
            continue
        p = T_r_h[host_of[i]][:3, :3] @ (f_host[i] / idist[i]) + T_r_h[host_of[i]][:3, 3]
        if p[2] >= 0.00001:
            ref[i] = np.linalg.norm(p)
    assert got.shape == (F,)
    assert np.array_equal(got < 0, exp < 0) and np.array_equal(exp < 0, ref < 0)
    assert all(got[i] == -1.0 for i in forced) and (got[has == 0] == -1.0).all() and (got >= 0).sum() > 250
    ok = exp >= 0
    assert np.abs(got[ok] - exp[ok]).max() < 1e-12 * np.abs(exp[ok]).max() and np.abs(ref[ok] - exp[ok]).max() < 1e-12 * np.abs(exp[ok]).max()


def test_stage_names_follow_the_reference_timers():
    """hso_stage_name: the reference's HSO_START_TIMER names (src/frame_handler_base.cpp:57-66) for the stages that have one."""
    from hso_b200 import _capi
    lib = _capi.load()
    names = [lib.hso_stage_name(i) for i in range(8)]
    assert names[:5] == [b"pyramid_creation", b"sparse_img_align", b"feature_align", b"pose_optimizer", b"reproject"]
    assert names[5] == b"depth_filter_update" and names[6] == b"feature_detection" and names[7] is None
    ref = open(os.path.join(ROOT, "tests", "golden", "reference_timer_names.txt")).read().split()
    assert all(n.decode() in ref for n in names[:5])


def test_compact_feature_layout_float32_px_is_exact_for_the_tracker():
    """hso_track_job::px32: the tracker only ever uses (float)(px * 2^-level) (src/CoarseTracker.cpp:437-441); rounding px to float32 first gives the
    same float at every level, because a power-of-two factor commutes with the rounding. xyz = f * dist is the reference's own expression (:292)."""
    import numpy as np
    from hso_b200 import Context
    rng = np.random.default_rng(0)
    px = np.stack([rng.uniform(0, 1280, 200000), rng.uniform(0, 1024, 200000)], axis=1)
    f = rng.normal(size=(200000, 3))
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    dist = rng.uniform(0.2, 50, 200000)
    dist[::17] = -1.0
    xyz, px32 = Context.compact_features(px, f, dist)
    ok = dist >= 0
    assert px32.dtype == np.float32 and px32.shape == (int(ok.sum()), 2) and xyz.shape == (int(ok.sum()), 3)
    assert np.array_equal(xyz, f[ok] * dist[ok, None])
    for level in range(0, 5):
        s = 1.0 / (1 << level)
        assert np.array_equal((px[ok] * s).astype(np.float32), (px32.astype(np.float64) * s).astype(np.float32))
