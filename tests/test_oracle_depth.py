"""CPU checks of the row-N3 oracle (oracle/oracle_depth.cpp). The reference ships no test for the depth filter, so the restatement is pinned by
properties: on a photoconsistent textured plane the epipolar search + KLT refinement must triangulate the TRUE depth, the Gaussian update must
move the estimate towards it and shrink the variance, invisible / degenerate seeds must take the reference's early exits."""
import collections

import numpy as np

import oracle_lib as O
from hso_b200 import synth
from hso_b200.api import Context


def _run(seed, cam, S, **kw):
    s = synth.make_depth_scene(seed, cam, S=S, **kw)
    oc = (O.orc_seed_obs * S).from_buffer_copy(bytes(Context.seed_obs(s["seeds"])))
    pyrs = [O.create_pyramid(im, 5)[0] for im in s["kf_imgs"]]
    cl, _ = O.create_pyramid(s["cur_img"], 5)
    sob = [O.sobel5(cl[l]) for l in range(3)]
    return s, O.depth_observe(s["cam"], s["T_cur_w"], s["T_f_w"], oc, s["px_error_angle"], pyrs, cl, sob)


def _true_depth(s, sd, plane=4.0):
    T = np.vstack([s["T_f_w"][sd["ref_pose"]], [0, 0, 0, 1]])
    Ti = np.linalg.inv(T)
    o, d = Ti[:3, 3], Ti[:3, :3] @ sd["f"]
    return (plane - o[2]) / d[2]


def test_line_stereo_triangulates_the_true_depth_and_updates_the_seed():
    s, out = _run(4, "icl", 400)
    S = len(s["seeds"])
    hist = collections.Counter(out[i].res for i in range(S))
    assert hist[1] > 0.5 * S and hist[-1] > 0 and hist[-4] > 0, hist   # successes and both kinds of rejection occur
    rel, closer = [], []
    for i, sd in enumerate(s["seeds"]):
        r = out[i]
        if r.res != 1:
            assert r.mu == np.float32(sd["mu"]) and r.sigma2 == np.float32(sd["sigma2"]) and list(r.epl_start) == [0, 0]
            continue
        lam = _true_depth(s, sd)
        rel.append(abs(r.z - lam) / lam)
        assert r.sigma2 <= np.float32(sd["sigma2"])                      # updateSeed never grows the variance (:535)
        closer.append(abs(r.mu - 1.0 / lam) <= abs(np.float32(sd["mu"]) - 1.0 / lam) + 1e-4)
        # the matched pixel lies on the epipolar segment's bounding box (a few px of slack for the KLT refinement)
        x0, x1 = sorted((r.epl_start[0], r.epl_end[0]))
        y0, y1 = sorted((r.epl_start[1], r.epl_end[1]))
        assert x0 - 6 <= r.px_cur[0] <= x1 + 6 and y0 - 6 <= r.px_cur[1] <= y1 + 6
    assert np.median(rel) < 5e-3 and np.quantile(rel, 0.9) < 3e-2
    assert np.mean(closer) > 0.9


def test_early_exits():
    s = synth.make_depth_scene(3, "icl", S=4)
    base = s["seeds"][1]
    seeds = [dict(base, mu=-0.3), dict(base, sigma2=float("nan")), dict(base, mu=1e-9, sigma2=1e-20)]
    oc = (O.orc_seed_obs * 3).from_buffer_copy(bytes(Context.seed_obs(seeds)))
    pyrs = [O.create_pyramid(im, 5)[0] for im in s["kf_imgs"]]
    cl, _ = O.create_pyramid(s["cur_img"], 5)
    sob = [O.sobel5(cl[l]) for l in range(3)]
    out = O.depth_observe(s["cam"], s["T_cur_w"], s["T_f_w"], oc, s["px_error_angle"], pyrs, cl, sob)
    assert out[0].is_update == 0 and out[0].res == 0            # behind the camera: not in view (:595-600)
    assert out[1].is_update == 1 and out[1].is_valid == 0       # NaN interval: isValid = false (:619), the match fails
    assert out[1].res != 1
    assert out[2].res != 1                                      # a point at 1e9 m: no usable epipolar segment / no match
