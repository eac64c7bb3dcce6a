"""One-off randomized parity sweep on a GPU box (not part of the test suite): runs the tracker per-evaluation parity check, the threshold
check and the N1 / N3 parity checks over more seeds, cameras and sizes than tests/ does and prints every failure."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_lib as O
import test_gpu_track as T
import test_gpu_reproject as R
import test_gpu_depth as D

O.load()
fails, runs = 0, 0


def attempt(name, fn):
    global fails, runs
    runs += 1
    try:
        fn()
    except Exception:
        fails += 1
        print("FAIL", name)
        traceback.print_exc(limit=3)


for seed in range(200, 200 + int(sys.argv[1]) if len(sys.argv) > 1 else 206):
    for cam, F in (("icl", 300 + 37 * (seed % 5)), ("euroc", 1500), ("tum_fov", 2500)):
        for ic in (False, True):
            def per_eval(seed=seed, cam=cam, F=F, ic=ic):
                p, ctx, tp, job, a0 = T._setup(O, seed, cam, F=F)
                res, traces = ctx.coarse_track_batch([job], inverse_comp=ic, trace_cap=256)
                T._check_trace(O, tp, traces[0], ic, 4)
                for lvl in (4, 3, 2, 1):
                    e = [t for t in traces[0] if t.level == lvl][0]
                    hu, ou, n = tp.select_robust(lvl, 4, np.array(e.T_eval[:]).reshape(3, 4), e.a_eval)
                    assert abs(hu - e.huber) <= 1e-5 * hu and abs(ou - e.outlier) <= 1e-5 * ou, (lvl, hu, e.huber, ou, e.outlier)
                ctx.close()
            attempt(f"track seed={seed} cam={cam} F={F} ic={ic}", per_eval)
    # N1 / N3: the parity tests of tests/ with their scene seed replaced
    r_run, d_run = R._run, D._run
    R._run = lambda oracle, cam, _seed, M, max_fts, _s=seed, **kw: r_run(oracle, cam, _s, M, max_fts, **kw)
    D._run = lambda oracle, cam, _seed, S, _s=seed, **kw: d_run(oracle, cam, _s, S, **kw)
    for cam, M, mf in (("icl", 900 + 100 * (seed % 7), 200), ("euroc", 1500, 400)):
        attempt(f"reproject seed={seed} cam={cam} M={M}", lambda cam=cam, M=M, mf=mf: R.test_reproject_match_parity(O, cam, M, mf))
    for cam, S in (("icl", 1500), ("euroc", 1000)):
        attempt(f"depth seed={seed} cam={cam}", lambda cam=cam, S=S: D.test_depth_observe_parity(O, cam, S, {}))
    R._run, D._run = r_run, d_run
print(f"{runs} runs, {fails} failures")
