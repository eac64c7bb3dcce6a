"""GPU test of the C++ host layer (hso_b200/host/hso_b200_host.hpp): compiled with g++ against libhso_b200.so, run on one synthetic
problem, compared with the ctypes path (same C-ABI underneath, so poses must agree to rounding of the 3x4 pose products)."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from hso_b200 import Context, make_cam, synth, _capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_layer_matches_ctypes_path(tmp_path):
    exe = tmp_path / "host_smoke"
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "cpp", "host_smoke.cpp"), "-o", str(exe),
                           f"-L{libdir}", "-lhso_b200", f"-Wl,-rpath,{libdir}"])
    p = synth.make_pair(91, "icl", F=600)
    c = p["cam"]
    F = len(p["dist"])
    has = (p["dist"] >= 0).astype(np.int32)
    idist = np.where(has == 1, 1.0 / np.abs(p["dist"]), 1.0)
    blob = tmp_path / "problem.bin"
    with open(blob, "wb") as f:
        f.write(struct.pack("<iii4d", c["width"], c["height"], F, c["fx"], c["fy"], c["cx"], c["cy"]))
        f.write(p["ref_img"].tobytes()); f.write(p["cur_img"].tobytes())
        f.write(np.ascontiguousarray(p["px"], np.float64).tobytes()); f.write(np.ascontiguousarray(p["f"], np.float64).tobytes())
        f.write(idist.astype(np.float64).tobytes()); f.write(has.tobytes())
    out = subprocess.run([str(exe), str(blob)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["threw"] == 1  # wrong image size throws like Frame::initFrame
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"]))
    ids, integral, _ = ctx.upload_frames([p["ref_img"], p["cur_img"]])
    a0 = float(np.float32(integral[1]) / np.float32(integral[0]))
    # makeDepthRef for self-hosted points: |f / idist| = dist (z > 1e-5 holds)
    dist = np.where(has == 1, np.linalg.norm(p["f"] * (1.0 / idist)[:, None], axis=1), -1.0)
    # the C++ layer's gather loop writes the compact feature layout (xyz = f * dist, float32 px, features without depth left out): same here, so
    # that both paths pick the same launch shape (it follows the feature count) and the iteration counts can be compared
    xyz, px32 = Context.compact_features(p["px"], p["f"], dist)
    res, _ = ctx.coarse_track_batch([dict(ref=ids[0], cur=ids[1], xyz=xyz, px32=px32, T_cur_ref=np.eye(4)[:3], exposure_rat=a0)])
    assert abs(r["integral"][0] - integral[0]) < 1e-4 and abs(r["integral"][1] - integral[1]) < 1e-4
    assert r["n_tracked"] == res[0]["n_tracked"]
    assert np.abs(np.array(r["T_track"]).reshape(3, 4) - res[0]["T_cur_ref"]).max() < 1e-9
    assert r["num_obs"] > 0.5 * has.sum() and r["launches"] >= 8
    assert np.abs(np.array(r["T"]).reshape(3, 4) - p["T_true"][:3]).max() < 5e-3
    # addImagesAndTrack (chunk-pipelined batch entry) reproduces the single-frame tracker exactly
    assert np.array_equal(np.array(r["T_batch"]).reshape(3, 4), np.array(r["T_track"]).reshape(3, 4))
    assert r["batch_iters"][0] == r["batch_iters"][1] == r["batch_iters"][2] == res[0]["n_iters"]
    # Reprojector::reprojectMap through the C++ layer == the same candidates through the ctypes path
    rp = r["reproj"]
    T_track = np.array(r["T_track"]).reshape(3, 4)
    idx = [i for i in range(F) if has[i]]
    cands = []
    for i in idx:
        ph = p["f"][i] / idist[i]
        cands.append(dict(p_host=ph, px_ref=p["px"][i], f_ref=p["f"][i], grad=np.array([1.0, 0.0]), depth_ref=1.0 / idist[i], host_pose=0, ref_pose=0,
                          ref_frame=ids[0], ref_level=0, ftr_type=0, pt_type=1 + i % 4, pt_ftr_type=i % 3, scale_patch=0, exposure_rat=1.0))
    cur2 = ctx.upload_frames([p["cur_img"]])[0][0]
    W_, H_ = c["width"], c["height"]
    cs = rp["cell_size"]
    grid = dict(cell_size=cs, n_cols=int(np.ceil(W_ / cs)), n_rows=int(np.ceil(H_ / cs)), max_fts=200, align_max_iter=10)
    out, summ = ctx.reproject_match(cur2, T_track, np.eye(4)[:3][None], Context.reproj_cands(cands), grid, np.array(rp["cell_order"], np.int32))
    assert (summ.n_matches, summ.n_trials, summ.n_in_frame) == (rp["n_matches"], rp["n_trials"], rp["n_in_frame"])
    assert rp["new_features"] == rp["n_matches"] > 50 and rp["failed"] == rp["n_trials"] - rp["n_matches"]
    first = [i for i in range(len(idx)) if out[i].order == 0][0]
    assert abs(out[first].px[0] - rp["first_px"][0]) < 1e-9 and abs(out[first].px[1] - rp["first_px"][1]) < 1e-9
    ctx.close()
