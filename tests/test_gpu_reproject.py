"""GPU parity, row N1: hso_reproject_match (reprojectPoint + grid cells + per-cell ordering + the three selection passes + the whole
Matcher::findMatchDirect incl. getWarpMatrixAffine) vs the CPU oracle.

Three layers, because the alignment itself is a float pipeline whose convergence can flip for a handful of candidates (see
test_gpu_align.py): (1) the fp64 geometry of every candidate (pixel, cell, warp matrix, search level) against the oracle, exact for the
integers; (2) the speculative findMatchDirect outcome of every candidate against the oracle's (<= 0.5 % flips); (3) the selection replayed
by the oracle's literal std::list walk on the device's own outcomes — tried / matched / creation order must agree EXACTLY."""
import ctypes as C

import numpy as np
import pytest

from hso_b200 import Context, HsoError, make_cam, synth

pytestmark = pytest.mark.gpu


def _run(oracle, cam, seed, M, max_fts, **kw):
    s = synth.make_reproject_scene(seed, cam, M=M, max_fts=max_fts, **kw)
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    arr = Context.reproj_cands(s["cands"], frame_ids=kf_ids)
    got, gsum = ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], arr, s["grid"], s["cell_order"], M=M)
    # oracle inputs: ref_frame indexes the list of keyframe pyramids
    oarr = Context.reproj_cands(s["cands"])
    oc = (oracle.orc_reproj_cand * max(M, 1)).from_buffer_copy(bytes(oarr))
    pyrs = [oracle.create_pyramid(im, 5)[0] for im in s["kf_imgs"]]
    cl, _ = oracle.create_pyramid(s["cur_img"], 5)
    sob = [oracle.sobel5(cl[l]) for l in range(3)]
    g = oracle.orc_reproj_grid(*[s["grid"][k] for k in ("cell_size", "n_cols", "n_rows", "max_fts", "align_max_iter")], 0)
    ctx.close()
    return s, got, gsum, oc, g, pyrs, cl, sob


@pytest.mark.parametrize("cam,M,max_fts", [("icl", 3000, 200), ("icl", 330, 200), ("icl", 500, 300), ("icl", 150, 200), ("euroc", 2500, 800),
                                           ("tum_fov", 1500, 200)])
def test_reproject_match_parity(oracle, cam, M, max_fts):
    s, got, gsum, oc, g, pyrs, cl, sob = _run(oracle, cam, 21, M, max_fts)
    spec, px_after = oracle.reproject_speculative(s["cam"], s["T_cur_w"], s["T_f_w"], oc, g, 2, pyrs, cl, sob)
    # (1) geometry
    inf_g = np.array([got[i].in_frame for i in range(M)])
    inf_o = np.array([spec[i].in_frame for i in range(M)])
    assert np.array_equal(inf_g, inf_o)
    assert 0.5 < inf_o.mean() <= 1.0
    assert [got[i].cell for i in range(M)] == [spec[i].cell for i in range(M)]
    elig = [i for i in range(M) if spec[i].in_frame and oc[i].pt_type != 0]
    n_job = 0
    for i in elig:
        Ao = np.array(spec[i].A_cur_ref[:])
        Ag = np.array(got[i].A_cur_ref[:])
        if np.any(Ao != 0):
            n_job += 1
            assert np.allclose(Ag, Ao, rtol=1e-9, atol=1e-9), (i, Ag, Ao)  # float in/out of undistortPoints is replicated, the rest is fp64
            assert got[i].search_level == spec[i].search_level
        else:
            assert not np.any(Ag != 0)  # getCloseViewObs failed or the reference pixel is too close to the border: no warp, no alignment
    assert n_job > 0.8 * len(elig)
    # (2) speculative outcome of findMatchDirect
    ok_g = np.array([got[i].align_ok for i in elig])
    ok_o = np.array([spec[i].align_ok for i in elig])
    assert 0.3 < ok_o.mean() < 0.99
    assert (ok_g != ok_o).mean() <= 0.005, (ok_g != ok_o).sum()
    # (3) the selection, replayed by the oracle on the device's outcomes
    io = (oracle.orc_reproj_result * M)()
    for i in range(M):
        io[i].in_frame, io[i].cell = got[i].in_frame, got[i].cell
    okd = np.array([got[i].align_ok for i in range(M)], np.uint8)
    osum = oracle.reproject_select(oc, okd, g, s["cell_order"], io)
    assert (gsum.used_cell_all, gsum.n_matches, gsum.n_trials, gsum.n_in_frame) == (osum.used_cell_all, osum.n_matches, osum.n_trials, osum.n_in_frame)
    assert [got[i].tried for i in range(M)] == [io[i].tried for i in range(M)]
    assert [got[i].matched for i in range(M)] == [io[i].matched for i in range(M)]
    assert [got[i].order for i in range(M)] == [io[i].order for i in range(M)]
    assert gsum.n_matches <= max_fts
    # pixels: untouched reprojection for untried candidates, the aligned pixel for tried ones
    for i in range(M):
        if not spec[i].in_frame:
            continue
        if got[i].tried:
            if ok_o[elig.index(i)] == got[i].align_ok and got[i].align_ok:
                assert np.hypot(got[i].px[0] - px_after[i, 0], got[i].px[1] - px_after[i, 1]) < 0.05, i
        else:
            assert abs(got[i].px[0] - spec[i].px[0]) < 1e-9 and abs(got[i].px[1] - spec[i].px[1]) < 1e-9


def test_reproject_match_end_to_end_equals_oracle_when_no_flip(oracle):
    """Whole call vs whole oracle (its own sequential walk with its own alignments). Seeds are scanned for a scene in which no
    speculative outcome differs; there every flag and the creation order must be identical."""
    for seed in range(30, 40):
        M, max_fts = 700, 200
        s, got, gsum, oc, g, pyrs, cl, sob = _run(oracle, "icl", seed, M, max_fts)
        spec, _ = oracle.reproject_speculative(s["cam"], s["T_cur_w"], s["T_f_w"], oc, g, 2, pyrs, cl, sob)
        if any(got[i].align_ok != spec[i].align_ok for i in range(M)):
            continue
        exp, esum = oracle.reproject_match(s["cam"], s["T_cur_w"], s["T_f_w"], oc, g, s["cell_order"], 2, pyrs, cl, sob)
        assert (gsum.n_matches, gsum.n_trials, gsum.n_in_frame, gsum.used_cell_all) == (esum.n_matches, esum.n_trials, esum.n_in_frame, esum.used_cell_all)
        for i in range(M):
            assert (got[i].tried, got[i].matched, got[i].order, got[i].cell) == (exp[i].tried, exp[i].matched, exp[i].order, exp[i].cell), i
            if got[i].matched:
                assert np.hypot(got[i].px[0] - exp[i].px[0], got[i].px[1] - exp[i].px[1]) < 0.05
                assert got[i].search_level == exp[i].search_level
        return
    pytest.fail("no flip-free scene in 10 seeds")


def test_reproject_edge_cases(oracle):
    s = synth.make_reproject_scene(5, "icl", M=40)
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"]), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    arr = Context.reproj_cands(s["cands"], frame_ids=kf_ids)
    # empty candidate list: nothing to do, reprojectCellAll branch
    out, summ = ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], arr, s["grid"], s["cell_order"], M=0)
    assert summ.n_matches == 0 and summ.n_in_frame == 0
    # a point behind the camera and a point with a bad pose index
    cands = [dict(s["cands"][0], p_host=np.array([0.0, 0.0, -3.0]))]
    out, summ = ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], Context.reproj_cands(cands, frame_ids=kf_ids), s["grid"], s["cell_order"], M=1)
    assert out[0].in_frame == 0 and out[0].tried == 0 and summ.n_in_frame == 0
    with pytest.raises(HsoError):
        ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], Context.reproj_cands([dict(s["cands"][0], host_pose=99)], frame_ids=kf_ids), s["grid"],
                            s["cell_order"], M=1)
    # cell_order must be a permutation
    bad = s["cell_order"].copy()
    bad[0] = bad[1]
    with pytest.raises(HsoError):
        ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], arr, s["grid"], bad)
    # max_fts = 0: the reference's loops break at once after the first cell / candidate
    g0 = dict(s["grid"], max_fts=0)
    out, summ = ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], arr, g0, s["cell_order"])
    oc = (oracle.orc_reproj_cand * 40).from_buffer_copy(bytes(Context.reproj_cands(s["cands"])))
    io = (oracle.orc_reproj_result * 40)()
    for i in range(40):
        io[i].in_frame, io[i].cell = out[i].in_frame, out[i].cell
    og = oracle.orc_reproj_grid(*[g0[k] for k in ("cell_size", "n_cols", "n_rows", "max_fts", "align_max_iter")], 0)
    osum = oracle.reproject_select(oc, np.array([out[i].align_ok for i in range(40)], np.uint8), og, s["cell_order"], io)
    assert [out[i].tried for i in range(40)] == [io[i].tried for i in range(40)] and summ.n_matches == osum.n_matches
    ctx.close()


def test_reproject_capacity_and_grid_errors():
    s = synth.make_reproject_scene(5, "icl", M=8)
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"]), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    big = Context.reproj_cands([s["cands"][0]] * 16385, frame_ids=kf_ids)
    with pytest.raises(HsoError) as e:
        ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], big, s["grid"], s["cell_order"])
    assert e.value.code == -4  # HSO_ERR_CAPACITY
    small_grid = dict(s["grid"], n_cols=2, n_rows=2)    # does not cover the image
    with pytest.raises(HsoError):
        ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], Context.reproj_cands(s["cands"], frame_ids=kf_ids), small_grid, np.arange(4, dtype=np.int32))
    ctx.close()


def test_selection_kernel_random_grids_incl_first_cell_third_pass(oracle):
    """The selection kernel alone (hso_reproject_select_only) against the oracle's std::list walk on 400 random grids, every 4th one with
    cell_order[0] holding >= 3 alignable candidates while the 3rd pass runs (the 2nd pass never visits that cell: src/reprojector.cpp:278-288).
    tried / matched / creation order / n_matches / n_trials (TYPE_DELETED entries are counted, :361-367) must agree exactly."""
    from test_oracle_reproject import random_selection_case
    from hso_b200 import _capi as K
    s = synth.make_reproject_scene(5, "icl", M=8)
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"]))
    rng = np.random.default_rng(77)
    multi = 0
    for trial in range(400):
        cands, in_frame, cell, ok, grid, order = random_selection_case(rng, force_pass3_in_first_cell=(trial % 4 == 0))
        M = len(in_frame)
        io = (oracle.orc_reproj_result * M)()
        for i in range(M):
            io[i].in_frame, io[i].cell = int(in_frame[i]), int(cell[i])
        osum = oracle.reproject_select(cands, ok, grid, order, io)
        kc = (K.hso_reproj_cand * M)()
        for i in range(M):
            kc[i].pt_type, kc[i].pt_ftr_type = cands[i].pt_type, cands[i].pt_ftr_type
        g = dict(cell_size=grid.cell_size, n_cols=grid.n_cols, n_rows=grid.n_rows, max_fts=grid.max_fts)
        got, gsum = ctx.reproject_select_only(kc, in_frame, cell, ok, g, order)
        assert (gsum.used_cell_all, gsum.n_matches, gsum.n_trials, gsum.n_in_frame) == (0, osum.n_matches, osum.n_trials, int(in_frame.sum())), trial
        assert [got[i].tried for i in range(M)] == [io[i].tried for i in range(M)], trial
        assert [got[i].matched for i in range(M)] == [io[i].matched for i in range(M)], trial
        assert [got[i].order for i in range(M)] == [io[i].order for i in range(M)], trial
        multi += sum(1 for i in range(M) if in_frame[i] and cell[i] == order[0] and io[i].matched) >= 3
    assert multi >= 50
    # the reprojectCellAll branch with deleted points in the list
    for trial in range(50):
        cands, in_frame, cell, ok, grid, order = random_selection_case(rng)
        M = len(in_frame)
        grid.max_fts = int(in_frame.sum())  # n_in_frame < max_fts + 50
        io = (oracle.orc_reproj_result * M)()
        for i in range(M):
            io[i].in_frame, io[i].cell = int(in_frame[i]), int(cell[i])
        osum = oracle.reproject_select(cands, ok, grid, order, io)
        kc = (K.hso_reproj_cand * M)()
        for i in range(M):
            kc[i].pt_type, kc[i].pt_ftr_type = cands[i].pt_type, cands[i].pt_ftr_type
        got, gsum = ctx.reproject_select_only(kc, in_frame, cell, ok, dict(cell_size=20, n_cols=grid.n_cols, n_rows=grid.n_rows, max_fts=grid.max_fts), order)
        assert (gsum.used_cell_all, gsum.n_matches, gsum.n_trials) == (1, osum.n_matches, osum.n_trials), trial
        assert [(got[i].tried, got[i].matched, got[i].order) for i in range(M)] == [(io[i].tried, io[i].matched, io[i].order) for i in range(M)], trial
    ctx.close()


# ---- a13b: the seed stage of Reprojector::reprojectMap (src/reprojector.cpp:309-328) + Matcher::findMatchSeed (src/matcher.cpp:442-518) ------------
def _seed_run(oracle, cam, seed, S, max_fts, gain, n_in=0):
    s = synth.make_seed_reproject_scene(seed, cam, S=S, max_fts=max_fts, gain=gain)
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    got, gsum = ctx.reproject_seeds(cur_id, s["T_cur_w"], s["T_f_w"], Context.seed_obs(s["seeds"], frame_ids=kf_ids), s["grid"], s["cell_order"], n_matches_in=n_in)
    oc = (oracle.orc_seed_obs * S).from_buffer_copy(bytes(Context.seed_obs(s["seeds"])))
    pyrs = [oracle.create_pyramid(im, 5)[0] for im in s["kf_imgs"]]
    cl, _ = oracle.create_pyramid(s["cur_img"], 5)
    sob = [oracle.sobel5(cl[l]) for l in range(3)]
    g = oracle.orc_reproj_grid(*[s["grid"][k] for k in ("cell_size", "n_cols", "n_rows", "max_fts", "align_max_iter")], 0)
    ctx.close()
    return s, got, gsum, oc, g, pyrs, cl, sob


@pytest.mark.parametrize("cam,S,max_fts,gain,n_in", [("icl", 2000, 200, 1.0, 0), ("icl", 800, 200, 1.3, 60), ("euroc", 1200, 120, 1.0, 90), ("tum_fov", 900, 200, 1.3, 0),
                                                     ("icl", 500, 30, 1.0, 28)])
def test_reproject_seeds_parity(oracle, cam, S, max_fts, gain, n_in):
    """Three layers like row N1: (1) fp64 geometry of reprojectorSeed + the head of findMatchSeed, exact for the integers; (2) speculative
    findMatchSeed outcome per seed (<= 0.5 % flips); (3) the per-cell sigma2 ordering + first-accepted-seed walk replayed by the oracle on the
    device's own outcomes — tried / matched / order / counters exactly."""
    s, got, gsum, oc, g, pyrs, cl, sob = _seed_run(oracle, cam, 61, S, max_fts, gain, n_in)
    spec, px_after = oracle.reproject_seeds_speculative(s["cam"], s["T_cur_w"], s["T_f_w"], oc, g, 2, pyrs, cl, sob)
    assert [got[i].in_frame for i in range(S)] == [spec[i].in_frame for i in range(S)]
    assert [got[i].cell for i in range(S)] == [spec[i].cell for i in range(S)]
    inf = [i for i in range(S) if spec[i].in_frame]
    assert len(inf) > 0.5 * S
    n_job = 0
    for i in inf:
        Ao, Ag = np.array(spec[i].A_cur_ref[:]), np.array(got[i].A_cur_ref[:])
        if np.any(Ao != 0):
            n_job += 1
            assert np.allclose(Ag, Ao, rtol=1e-9, atol=1e-9) and got[i].search_level == spec[i].search_level, i
        else:
            assert not np.any(Ag != 0)
    assert n_job > 0.8 * len(inf)
    ok_g = np.array([got[i].align_ok for i in inf])
    ok_o = np.array([spec[i].align_ok for i in inf])
    assert 0.05 < ok_o.mean() < 0.99 and (ok_g != ok_o).mean() <= 0.005, (ok_g != ok_o).sum()
    io = (oracle.orc_reproj_result * S)()
    for i in range(S):
        io[i].in_frame, io[i].cell = got[i].in_frame, got[i].cell
    osum = oracle.seed_select(oc, np.array([got[i].align_ok for i in range(S)], np.uint8), g, s["cell_order"], n_in, io)
    assert (gsum.n_matches, gsum.n_trials, gsum.n_in_frame, gsum.used_cell_all) == (osum.n_matches, osum.n_trials, osum.n_in_frame, 0)
    assert [(got[i].tried, got[i].matched, got[i].order) for i in range(S)] == [(io[i].tried, io[i].matched, io[i].order) for i in range(S)]
    assert gsum.n_matches <= max(max_fts, n_in + 1)
    for i in inf:
        if got[i].tried and got[i].align_ok and spec[i].align_ok:
            assert np.hypot(got[i].px[0] - px_after[i, 0], got[i].px[1] - px_after[i, 1]) < 0.05
        elif not got[i].tried:
            assert abs(got[i].px[0] - spec[i].px[0]) < 1e-9 and abs(got[i].px[1] - spec[i].px[1]) < 1e-9


def test_reproject_seeds_end_to_end_and_edges(oracle):
    """Whole call vs whole oracle on a flip-free scene; S = 0; n_matches_in already at maxFts (the loop still visits the first cell)."""
    for seed in range(70, 80):
        S = 700
        s, got, gsum, oc, g, pyrs, cl, sob = _seed_run(oracle, "icl", seed, S, 200, 1.0, 20)
        spec, _ = oracle.reproject_seeds_speculative(s["cam"], s["T_cur_w"], s["T_f_w"], oc, g, 2, pyrs, cl, sob)
        if any(got[i].align_ok != spec[i].align_ok for i in range(S)):
            continue
        exp, esum = oracle.reproject_seeds(s["cam"], s["T_cur_w"], s["T_f_w"], oc, g, s["cell_order"], 20, 2, pyrs, cl, sob)
        assert (gsum.n_matches, gsum.n_trials, gsum.n_in_frame) == (esum.n_matches, esum.n_trials, esum.n_in_frame)
        assert [(got[i].tried, got[i].matched, got[i].order) for i in range(S)] == [(exp[i].tried, exp[i].matched, exp[i].order) for i in range(S)]
        break
    else:
        pytest.fail("no flip-free scene in 10 seeds")
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"]), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    arr = Context.seed_obs(s["seeds"], frame_ids=kf_ids)
    out, summ = ctx.reproject_seeds(cur_id, s["T_cur_w"], s["T_f_w"], arr, s["grid"], s["cell_order"], n_matches_in=7, S=0)
    assert summ.n_matches == 7 and summ.n_trials == 0
    out, summ = ctx.reproject_seeds(cur_id, s["T_cur_w"], s["T_f_w"], arr, s["grid"], s["cell_order"], n_matches_in=200)
    first = int(s["cell_order"][0])
    assert all(out[i].tried == 0 for i in range(S) if out[i].cell != first) and summ.n_matches in (200, 201)
    with pytest.raises(HsoError):
        ctx.reproject_seeds(cur_id, s["T_cur_w"], s["T_f_w"], Context.seed_obs([dict(s["seeds"][0], ref_pose=99)], frame_ids=kf_ids), s["grid"], s["cell_order"])
    ctx.close()


def test_stage_timers_of_a_reprojection_call(oracle):
    """"reproject" times the whole call, "feature_align" the alignment kernel nested inside it (src/reprojector.cpp:96,259,330)."""
    s = synth.make_reproject_scene(5, "icl", M=300)
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"]), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], Context.reproj_cands(s["cands"], frame_ids=kf_ids), s["grid"], s["cell_order"])
    ms_r, n_r = ctx.stage_time_ms(4)
    ms_a, n_a = ctx.stage_time_ms(2)
    ms_p, n_p = ctx.stage_time_ms(0)
    assert n_r == 1 and n_a == 1 and n_p == 2 and 0 < ms_a < ms_r
    ctx.close()
