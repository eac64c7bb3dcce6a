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
    assert C.sizeof(K.hso_track_job) == 16 + 3 * 8 + 96 + 8
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
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
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
