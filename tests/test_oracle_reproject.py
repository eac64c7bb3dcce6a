"""CPU checks of the row-N1 oracle (oracle/oracle_reproject.cpp): cam2world incl. the cv::undistortPoints branch against cv2 4.13 golden
vectors (bit-exact), warp::getWarpMatrixAffine against the analytic plane-homography Jacobian, and the selection walk against an
independent brute-force statement of the reference's three passes in Python."""
import ctypes as C
import os

import numpy as np

import oracle_lib as O
from hso_b200 import synth

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv_golden2.npz"))


def test_cam2world_radtan_matches_cv_undistort_points():
    K, d = GOLD["undist_K"], GOLD["undist_d"]
    cam = dict(model=0, width=752, height=480, fx=K[0, 0], fy=K[1, 1], cx=K[0, 2], cy=K[1, 2], d=d)
    for uv, xy in zip(GOLD["undist_uv"], GOLD["undist_xy"]):
        b = O.cam2world(cam, uv[0], uv[1])
        x, y = float(xy[0]), float(xy[1])  # cv2's float32 output, as the reference reads it back
        n = np.sqrt(x * x + y * y + 1.0)
        assert abs(b[0] - x / n) <= 1e-16 and abs(b[1] - y / n) <= 1e-16 and abs(b[2] - 1.0 / n) <= 1e-16


def test_cam2world_inverts_world2cam_all_models():
    rng = np.random.default_rng(3)
    for name in ("icl", "euroc", "tum_fov"):
        cam = synth.CAMS[name]
        for _ in range(50):
            P = np.array([rng.uniform(-1, 1), rng.uniform(-0.7, 0.7), rng.uniform(1.5, 6)])
            px = synth.world2cam(cam, P)[0]
            b = O.cam2world(cam, px[0], px[1])
            tol = 2e-4 if name == "euroc" else 1e-9  # radtan: float in/out and a 5-step fixed point (src/camera.cpp:74-84)
            assert np.allclose(b, P / np.linalg.norm(P), atol=tol), (name, b, P / np.linalg.norm(P))
            assert abs(np.linalg.norm(b) - 1) < 1e-12


def test_warp_matrix_is_the_plane_homography_jacobian():
    cam = synth.CAMS["icl"]
    K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1.0]])
    rng = np.random.default_rng(5)
    depth = 4.0
    for lvl in (0, 1, 2):
        T = synth.se3_exp(np.concatenate([rng.normal(0, 0.05, 3), rng.normal(0, 0.02, 3)]))
        Hm = synth.homography(K, T, depth)
        px = np.array([rng.uniform(100, 540), rng.uniform(100, 380)])
        ray = np.array([(px[0] - cam["cx"]) / cam["fx"], (px[1] - cam["cy"]) / cam["fy"], 1.0])
        f = ray / np.linalg.norm(ray)
        A = O.get_warp_matrix_affine(cam, px, f, depth / f[2], T[:3], lvl)

        def w(p):
            q = Hm @ np.array([p[0], p[1], 1.0])
            return q[:2] / q[2]
        s = 5.0 * (1 << lvl)
        # matcher.cpp:69-70 divides by halfpatch_size only: at level l the columns are (1<<l) x the level-0 Jacobian
        Afd = np.stack([(w(px + [s, 0]) - w(px)) / 5.0, (w(px + [0, s]) - w(px)) / 5.0], axis=1)
        assert np.allclose(A, Afd, atol=1e-9), (lvl, A, Afd)


def _brute_force_selection(cands, in_frame, cell, ok, grid, cell_order):
    """Independent statement of src/reprojector.cpp:253-303,351-424,545-615 with Python lists."""
    M = len(cands)
    n_cells = grid.n_cols * grid.n_rows
    cells = [[] for _ in range(n_cells)]
    allp = []
    for i in range(M):
        if in_frame[i]:
            cells[cell[i]].append(i)
            allp.append(i)
    tried, matched, order = np.zeros(M, int), np.zeros(M, int), -np.ones(M, int)
    st = dict(n=0, o=0, trials=0)

    def hit(i):
        matched[i] = 1
        order[i] = st["o"]
        st["o"] += 1
    if len(allp) < grid.max_fts + 50:
        for i in allp:
            st["trials"] += 1  # ++n_trials_ precedes the TYPE_DELETED test (src/reprojector.cpp:553-559)
            if cands[i].pt_type == 0:
                continue
            tried[i] = 1
            if ok[i]:
                hit(i)
                st["n"] += 1
                if st["n"] >= grid.max_fts:
                    break
        return tried, matched, order, st["n"], 1, st["trials"]

    def reproject_cell(lst, is_2nd, is_3rd):
        if not lst:
            return False
        if not is_2nd:
            lst.sort(key=lambda i: (-cands[i].pt_type, -cands[i].pt_ftr_type))  # Python's sort is stable, like std::list::sort
        succ = 0
        while lst:
            i = lst.pop(0)
            st["trials"] += 1  # src/reprojector.cpp:361-367
            if cands[i].pt_type == 0:
                continue
            tried[i] = 1
            if not ok[i]:
                continue
            hit(i)
            if not is_3rd:
                return True
            succ += 1
            st["n"] += 1
            if succ >= 3 or st["n"] >= grid.max_fts:
                return True
        return False
    for i in range(n_cells):
        if reproject_cell(cells[cell_order[i]], False, False):
            st["n"] += 1
        if st["n"] >= grid.max_fts:
            break
    if st["n"] < grid.max_fts:
        for i in range(n_cells - 1, 0, -1):
            if reproject_cell(cells[cell_order[i]], True, False):
                st["n"] += 1
            if st["n"] >= grid.max_fts:
                break
    if st["n"] < grid.max_fts:
        for i in range(n_cells):
            reproject_cell(cells[cell_order[i]], True, True)
            if st["n"] >= grid.max_fts:
                break
    return tried, matched, order, st["n"], 0, st["trials"]


def test_selection_walk_matches_brute_force():
    rng = np.random.default_rng(11)
    for trial, (M, n_cols, n_rows, max_fts, p_ok) in enumerate([(600, 10, 8, 60, 0.6), (300, 10, 8, 100, 0.3), (90, 6, 5, 60, 0.7),
                                                               (500, 12, 9, 400, 0.8), (260, 8, 8, 150, 0.5), (1, 4, 4, 10, 1.0)]):
        cands = (O.orc_reproj_cand * M)()
        io = (O.orc_reproj_result * M)()
        n_cells = n_cols * n_rows
        for i in range(M):
            cands[i].pt_type = int(rng.choice([0, 1, 2, 3, 4], p=[0.05, 0.15, 0.2, 0.3, 0.3]))
            cands[i].pt_ftr_type = int(rng.integers(0, 3))
            io[i].in_frame = int(rng.uniform() < 0.85)
            io[i].cell = int(rng.integers(0, n_cells)) if io[i].in_frame else -1
        ok = (rng.uniform(size=M) < p_ok).astype(np.uint8)
        grid = O.orc_reproj_grid(20, n_cols, n_rows, max_fts, 10, 0)
        cell_order = rng.permutation(n_cells).astype(np.int32)
        exp = _brute_force_selection(cands, [io[i].in_frame for i in range(M)], [io[i].cell for i in range(M)], ok, grid, cell_order)
        summ = O.reproject_select(cands, ok, grid, cell_order, io)
        assert summ.used_cell_all == exp[4] and summ.n_matches == exp[3] and summ.n_trials == exp[5], trial
        assert [io[i].tried for i in range(M)] == list(exp[0]), trial
        assert [io[i].matched for i in range(M)] == list(exp[1]), trial
        assert [io[i].order for i in range(M)] == list(exp[2]), trial
        assert summ.n_matches <= max_fts and sorted(o for o in exp[2] if o >= 0) == list(range(summ.n_matches))


def random_selection_case(rng, force_pass3_in_first_cell=False):
    """A random grid / candidate set for the three selection passes. With force_pass3_in_first_cell the configuration the round-1 advisor found
    (pass 3 reached, cell_order[0] holding >= 3 alignable candidates: the 2nd pass never visits that cell, so its 2nd..4th successes are all
    taken by the 3rd pass) is constructed explicitly."""
    n_cols, n_rows = int(rng.integers(2, 9)), int(rng.integers(2, 8))
    nc = n_cols * n_rows
    maxf = int(rng.integers(0, 60))
    M = int(rng.integers(maxf + 60, maxf + 400))
    p_ok = float(rng.choice([0.05, 0.2, 0.5, 0.9]))
    cands = (O.orc_reproj_cand * M)()
    in_frame = (rng.uniform(size=M) < 0.9).astype(np.int32)
    cell = np.where(in_frame > 0, rng.integers(0, nc, M), -1).astype(np.int32)
    ok = (rng.uniform(size=M) < p_ok).astype(np.uint8)
    order = rng.permutation(nc).astype(np.int32)
    for i in range(M):
        cands[i].pt_type = int(rng.choice([0, 1, 2, 3, 4], p=[0.08, 0.12, 0.2, 0.3, 0.3]))
        cands[i].pt_ftr_type = int(rng.integers(0, 3))
    if force_pass3_in_first_cell:
        maxf = M  # never reached: all three passes run to the end
        k = int(rng.integers(3, 7))
        idx = rng.choice(M, k, replace=False)
        in_frame[idx], cell[idx], ok[idx] = 1, order[0], 1
        for i in idx:
            cands[int(i)].pt_type = int(rng.integers(1, 5))
    if int(in_frame.sum()) < maxf + 50:  # stay in the three-pass branch
        maxf = max(0, int(in_frame.sum()) - 50)
    grid = O.orc_reproj_grid(20, n_cols, n_rows, maxf, 10, 0)
    return cands, in_frame, cell, ok, grid, order


def test_selection_walk_random_grids_incl_first_cell_third_pass():
    """1200 random grids (every 4th one the forced cell_order[0] case): the oracle's std::list walk against the brute-force statement, all
    flags, the creation order and both counters."""
    rng = np.random.default_rng(2024)
    seen_first_cell_multi = 0
    for trial in range(1200):
        cands, in_frame, cell, ok, grid, order = random_selection_case(rng, force_pass3_in_first_cell=(trial % 4 == 0))
        M = len(in_frame)
        io = (O.orc_reproj_result * M)()
        for i in range(M):
            io[i].in_frame, io[i].cell = int(in_frame[i]), int(cell[i])
        exp = _brute_force_selection(cands, list(in_frame), list(cell), ok, grid, order)
        summ = O.reproject_select(cands, ok, grid, order, io)
        assert exp[4] == 0 and summ.used_cell_all == 0
        assert (summ.n_matches, summ.n_trials) == (exp[3], exp[5]), trial
        assert [io[i].tried for i in range(M)] == list(exp[0]) and [io[i].matched for i in range(M)] == list(exp[1]), trial
        assert [io[i].order for i in range(M)] == list(exp[2]), trial
        assert sorted(o for o in exp[2] if o >= 0) == list(range(summ.n_matches)), trial  # a permutation: no collision, nothing out of range
        if sum(1 for i in range(M) if in_frame[i] and cell[i] == order[0] and exp[1][i]) >= 3:
            seen_first_cell_multi += 1
    assert seen_first_cell_multi >= 100
