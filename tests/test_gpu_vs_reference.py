"""GPU parity against the REFERENCE ITSELF (oracle/_ref/libhso_ref.so: the reference's own sources compiled unmodified, see oracle/ref_wrap.cpp)
through the C-ABI — the same checks tests/test_gpu_track.py / test_gpu_frame.py / test_gpu_pose.py / test_gpu_reproject.py run against the
oracle restatement, with the reference's functions as the checker: hso::Frame's constructor, CoarseTracker::{precomputeReferencePatches,
computeResiduals, computeGS, selectRobustFunctionLevel, run}, Matcher::findMatchDirect, pose_optimizer::optimizeLevenbergMarquardt3rd.
The library is built where /root/reference exists and travels to the GPU box as a prebuilt file."""
import numpy as np
import pytest

import ref_lib as R
from hso_b200 import Context, make_cam, synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref/libhso_ref.so not built")]

REL = 1e-4  # north_star: pose increment per iteration to 1e-4 relative


class _RefProblem:
    """Adapter: the reference's own per-evaluation functions behind the interface test_gpu_track._check_trace expects."""

    def __init__(self, ref_frame, cur_frame):
        self.rf, self.cf = ref_frame, cur_frame

    def eval(self, level, max_level, T, a, huber, outlier, inverse_comp=False):
        return R.track_eval(self.rf, self.cf, level, max_level, T, a, huber, outlier, inverse_comp=inverse_comp)


def _setup(seed, cam, F):
    p = synth.make_pair(seed, cam, F=F)
    c = p["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)))
    ids, integral, gm = ctx.upload_frames([p["ref_img"], p["cur_img"]])
    rf, cf = R.Frame(c, p["ref_img"]), R.Frame(c, p["cur_img"])
    rf.set_track_features(p["px"], p["f"], p["dist"])
    return p, ctx, ids, integral, gm, rf, cf


@pytest.mark.parametrize("cam", ["icl", "euroc", "tum_fov"])
def test_frame_construction_vs_reference_frame(cam):
    """hso_frame_upload vs new hso::Frame(cam, img): pyramid bytes and Sobel images bit for bit, the two statistics to the float running sum's
    own accuracy (the device sums exactly; the reference accumulates ~2.7e5 values into one float)."""
    c = synth.CAMS[cam]
    img = synth.texture(np.random.default_rng(3), c["width"], c["height"])
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), materialize_sobel=True)
    ids, integral, gm = ctx.upload_frames([img])
    fr = R.Frame(c, img)
    lv = fr.levels()
    for l in range(5):
        assert np.array_equal(ctx.download_level(ids[0], l), lv[l]), (cam, l)
    for l in range(3):
        gx, gy = ctx.download_sobel(ids[0], l)
        rx, ry = fr.sobel(l)
        assert np.array_equal(gx, rx) and np.array_equal(gy, ry)
    ri, rg = fr.stats()
    assert abs(integral[0] - ri) <= 5e-5 * ri and abs(gm[0] - rg) <= 2.5e-4 * rg
    fr.close()
    ctx.close()


@pytest.mark.parametrize("ic", [False, True])
@pytest.mark.parametrize("cam,F", [("icl", 3000), ("euroc", 2000), ("tum_fov", 1500)])
def test_tracker_trace_vs_reference_functions(oracle, cam, F, ic):
    """Every evaluation the device traced, replayed by the reference's own computeResiduals + computeGS at the same state (H, b, energy, term
    counts), every step against the damped solve of the reference-built system (<= 1e-4 relative), thresholds against the reference's
    selectRobustFunctionLevel."""
    from test_gpu_track import _check_trace
    p, ctx, ids, integral, gm, rf, cf = _setup(300 + F, cam, F)
    a0 = float(np.float32(cf.stats()[0]) / np.float32(rf.stats()[0]))
    job = dict(ref=ids[0], cur=ids[1], px=p["px"], f=p["f"], dist=p["dist"], T_cur_ref=np.eye(4)[:3], exposure_rat=a0)
    res, traces = ctx.coarse_track_batch([job], inverse_comp=ic, trace_cap=256)
    worst = _check_trace(oracle, _RefProblem(rf, cf), traces[0], ic, 4)
    assert worst <= REL
    seen = set()
    for e in traces[0]:
        if e.iter == -1 and e.level not in seen:
            seen.add(e.level)
            hu, ou = R.track_select_robust(rf, cf, e.level, 4, np.array(e.T_eval[:]).reshape(3, 4), e.a_eval)
            assert abs(hu - e.huber) <= 1e-5 * hu + 5e-5 and abs(ou - e.outlier) <= 1e-5 * ou + 1.5e-4
    assert seen == {4, 3, 2, 1}
    rf.close(); cf.close(); ctx.close()


@pytest.mark.parametrize("ic", [False, True])
@pytest.mark.parametrize("cam,F,min_level,n_iter", [("icl", 3000, 1, 50), ("icl", 1500, 1, 50), ("icl", 1000, 0, 15), ("euroc", 2000, 1, 50)])
def test_tracker_full_run_vs_reference_run(cam, F, min_level, n_iter, ic):
    """hso_coarse_track vs CoarseTracker::run of the reference on the same frames (exposure ratio formed from the frames' statistics on both
    sides): final pose, exposure ratio, return value. (The wide FOV camera, whose image corners lie outside the lens model, is covered
    evaluation by evaluation in test_tracker_trace_vs_reference_functions.)"""
    p, ctx, ids, integral, gm, rf, cf = _setup(500 + F + min_level, cam, F)
    job = dict(ref=ids[0], cur=ids[1], px=p["px"], f=p["f"], dist=p["dist"], T_cur_ref=np.eye(4)[:3], exposure_rat=-1.0)
    res, _ = ctx.coarse_track_batch([job], inverse_comp=ic, min_level=min_level, n_iter=n_iter)
    rr = R.coarse_track(rf, cf, np.eye(4)[:3], inverse_comp=ic, min_level=min_level, n_iter=n_iter)
    tol = 2e-4 if cam == "icl" else 5e-4  # radtan: the interpolation of a distorted render leaves a slightly shallower optimum
    assert np.abs(res[0]["T_cur_ref"] - rr["T_cur_ref"]).max() < tol
    assert abs(res[0]["exposure_rat"] - rr["exposure_rat"]) < tol
    assert abs(res[0]["n_tracked"] - rr["n_tracked"]) <= 2
    rf.close(); cf.close(); ctx.close()


@pytest.mark.parametrize("F,K,seed", [(400, 8, 1), (5000, 8, 3), (60, 3, 2)])
def test_pose_optimizer_vs_reference(F, K, seed):
    """hso_pose_optimize vs pose_optimizer::optimizeLevenbergMarquardt3rd of the reference on real Frame / Feature / Point objects."""
    p = synth.make_pose_problem(seed, "icl", F=F, K=K)
    c = synth.CAMS["icl"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"]))
    g = ctx.pose_optimize_batch([p])[0]
    r = R.pose_optimize(c, p)
    assert np.abs(g["T_f_w"] - r["T_f_w"]).max() < 1e-7
    assert abs(g["estimated_scale"] - r["estimated_scale"]) <= 1e-6 * abs(r["estimated_scale"])
    assert abs(g["error_init"] - r["error_init"]) <= 1e-5 * r["error_init"] and abs(g["error_final"] - r["error_final"]) <= 1e-5 * r["error_final"]
    assert int((g["outlier"] != r["outlier"]).sum()) <= 1
    ctx.close()


@pytest.mark.parametrize("cam,M", [("icl", 1500), ("euroc", 1000)])
def test_find_match_direct_vs_reference(oracle, cam, M):
    """hso_reproject_match's speculative per-candidate outcome vs Matcher::findMatchDirect of the reference on real Point / Feature / Frame
    objects: warp matrix, search level, return value (<= 0.5 % flips of the float alignment), final pixel."""
    s = synth.make_reproject_scene(41, cam, M=M, max_fts=200, gain=1.3)
    c = s["cam"]
    for cd in s["cands"]:
        cd["scale_patch"] = 1  # keyframe gap 1 < 4 and |128 * 1.3 - 128| > 30: the reference's own rule says "scale"
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    got, gsum = ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], Context.reproj_cands(s["cands"], frame_ids=kf_ids), s["grid"], s["cell_order"], M=M)
    kfs = [R.Frame(c, im, T, exposure_time=1.0, keyframe_id=1) for im, T in zip(s["kf_imgs"], s["T_f_w"])]
    cur = R.Frame(c, s["cur_img"], s["T_cur_w"], exposure_time=1.3, keyframe_id=2)
    oc = (oracle.orc_reproj_cand * M).from_buffer_copy(bytes(Context.reproj_cands(s["cands"])))
    # the reprojected pixel findMatchDirect starts from: Reprojector::reprojectPoint through the reference's own camera + Sophus
    px0 = np.zeros((M, 2))
    for i, cd in enumerate(s["cands"]):
        Tth = R.se3_mul(s["T_cur_w"], R.se3_inverse(s["T_f_w"][cd["host_pose"]]))
        P = Tth[:, :3] @ np.asarray(cd["p_host"]) + Tth[:, 3]
        px0[i] = R.world2cam(c, P) if P[2] > 1e-5 else 0
    ok, px, sl, A, _ = R.find_match_batch(cur, kfs, oc, px0)
    n_job = flips = 0
    for i in range(M):
        if not got[i].in_frame or oc[i].pt_type == 0 or oc[i].ref_pose < 0:
            continue
        assert abs(got[i].px[0] - px0[i, 0]) < 1e-6 or got[i].tried  # same reprojection
        Ag = np.array(got[i].A_cur_ref[:]).reshape(2, 2)
        if not np.any(Ag != 0):
            assert ok[i] == 0
            continue
        n_job += 1
        assert np.allclose(Ag, A[i], rtol=1e-9, atol=1e-9) and got[i].search_level == sl[i], i
        flips += int(got[i].align_ok != ok[i])
    assert n_job > 0.5 * M and flips <= 0.005 * n_job, (flips, n_job)
    for k in kfs:
        k.close()
    cur.close()
    ctx.close()


@pytest.mark.parametrize("cam,S,gain", [("icl", 1200, 1.3), ("euroc", 800, 1.0)])
def test_find_match_seed_vs_reference(oracle, cam, S, gain):
    """hso_reproject_seeds' speculative per-seed outcome vs Matcher::findMatchSeed of the reference (a13b)."""
    s = synth.make_seed_reproject_scene(43, cam, S=S, gain=gain)
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    got, gsum = ctx.reproject_seeds(cur_id, s["T_cur_w"], s["T_f_w"], Context.seed_obs(s["seeds"], frame_ids=kf_ids), s["grid"], s["cell_order"])
    kfs = [R.Frame(c, im, T, exposure_time=1.0, keyframe_id=1) for im, T in zip(s["kf_imgs"], s["T_f_w"])]
    cur = R.Frame(c, s["cur_img"], s["T_cur_w"], exposure_time=gain, keyframe_id=9)
    oc = (oracle.orc_seed_obs * S).from_buffer_copy(bytes(Context.seed_obs(s["seeds"])))
    px0 = np.zeros((S, 2))
    for i, sd in enumerate(s["seeds"]):
        Tth = R.se3_mul(s["T_cur_w"], R.se3_inverse(s["T_f_w"][sd["ref_pose"]]))
        P = Tth[:, :3] @ (np.asarray(sd["f"]) * (1.0 / float(np.float32(sd["mu"])))) + Tth[:, 3]
        px0[i] = R.world2cam(c, P) if P[2] >= 0.001 else 0
    ok, px, sl, A = R.find_match_seed_batch(cur, kfs, oc, px0)
    n_job = flips = 0
    for i in range(S):
        if not got[i].in_frame:
            continue
        Ag = np.array(got[i].A_cur_ref[:]).reshape(2, 2)
        if not np.any(Ag != 0):
            assert ok[i] == 0
            continue
        n_job += 1
        assert np.allclose(Ag, A[i], rtol=1e-7, atol=1e-7) and got[i].search_level == sl[i], i
        flips += int(got[i].align_ok != ok[i])
    assert n_job > 0.5 * S and flips <= 0.005 * n_job, (flips, n_job)
    for k in kfs:
        k.close()
    cur.close()
    ctx.close()


@pytest.mark.parametrize("cam,S", [("icl", 2000), ("euroc", 1200), ("tum_fov", 1000)])
def test_depth_observe_vs_reference(oracle, cam, S):
    """hso_depth_observe vs DepthFilter::observeDepthRow of the reference itself (row N3): visibility / validity exact, the outcome of
    Matcher::doLineStereo with <= 1 % flips, and where both succeed the epipolar end points, the search level, the matched pixel and the updated seed."""
    s = synth.make_depth_scene(8, cam, S=S)
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    got = ctx.depth_observe(cur_id, s["T_cur_w"], s["T_f_w"], Context.seed_obs(s["seeds"], frame_ids=kf_ids), s["px_error_angle"], S=S)
    kfs = [R.Frame(c, im, T, exposure_time=1.0, keyframe_id=1) for im, T in zip(s["kf_imgs"], s["T_f_w"])]
    cur = R.Frame(c, s["cur_img"], s["T_cur_w"], exposure_time=1.0, keyframe_id=9)
    oc = (oracle.orc_seed_obs * S).from_buffer_copy(bytes(Context.seed_obs(s["seeds"])))
    ref = R.depth_observe(cur, kfs, oc, s["px_error_angle"])
    assert [got[i].is_update for i in range(S)] == [ref[i].is_update for i in range(S)]
    upd = [i for i in range(S) if ref[i].is_update]
    assert [got[i].is_valid for i in upd] == [ref[i].is_valid for i in upd]
    flips = sum(1 for i in upd if (got[i].res == 1) != (ref[i].res == 1) or (got[i].res != 1 and got[i].res != ref[i].res))
    assert flips <= max(1, 0.01 * len(upd)), (flips, len(upd))
    both = [i for i in upd if got[i].res == 1 and ref[i].res == 1]
    assert len(both) >= 30
    for i in both:
        assert list(got[i].epl_start) == list(ref[i].epl_start) and list(got[i].epl_end) == list(ref[i].epl_end) and got[i].search_level == ref[i].search_level
    dpx = np.array([np.hypot(got[i].px_cur[0] - ref[i].px_cur[0], got[i].px_cur[1] - ref[i].px_cur[1]) for i in both])
    dmu = np.array([abs(got[i].mu - ref[i].mu) / abs(ref[i].mu) for i in both])
    dsg = np.array([abs(got[i].sigma2 - ref[i].sigma2) / ref[i].sigma2 for i in both])
    assert np.median(dpx) < 1e-3 and np.quantile(dpx, 0.99) < 0.05
    assert np.median(dmu) < 1e-5 and np.quantile(dmu, 0.99) < 5e-3
    assert np.median(dsg) < 1e-4 and np.quantile(dsg, 0.99) < 5e-2
    for k in kfs:
        k.close()
    cur.close()
    ctx.close()


@pytest.mark.parametrize("cam,M,max_fts,seed", [("icl", 3000, 200, 21), ("euroc", 1500, 150, 23), ("icl", 180, 200, 25)])
def test_reproject_match_vs_reference(oracle, cam, M, max_fts, seed):
    """hso_reproject_match (k_reproject + k_align on every candidate + k_reproj_select) vs the grid stage of Reprojector::reprojectMap run through the
    reference's own reprojectPoint / reprojectCell / reprojectCellAll (row N1): in-frame flag and cell per point exact; tried / matched flags,
    creation order and both counters equal when no alignment outcome flips along the walk, else within the flip bound."""
    s = synth.make_reproject_scene(seed, cam, M=M, max_fts=max_fts)
    c = s["cam"]
    ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), materialize_sobel=True)
    kf_ids, _, _ = ctx.upload_frames(s["kf_imgs"])
    cur_id = ctx.upload_frames([s["cur_img"]])[0][0]
    got, gsum = ctx.reproject_match(cur_id, s["T_cur_w"], s["T_f_w"], Context.reproj_cands(s["cands"], frame_ids=kf_ids), s["grid"], s["cell_order"])
    kfs = [R.Frame(c, im, T, exposure_time=1.0, keyframe_id=1) for im, T in zip(s["kf_imgs"], s["T_f_w"])]
    cur = R.Frame(c, s["cur_img"], s["T_cur_w"], exposure_time=1.0, keyframe_id=2)
    oc = (oracle.orc_reproj_cand * M).from_buffer_copy(bytes(Context.reproj_cands(s["cands"])))
    g = oracle.orc_reproj_grid(*[s["grid"][k] for k in ("cell_size", "n_cols", "n_rows", "max_fts", "align_max_iter")], 0)
    ref, rsum = R.reproject_match(cur, kfs, oc, g, s["cell_order"])
    assert rsum.n_matches >= 0
    assert [got[i].in_frame for i in range(M)] == [ref[i].in_frame for i in range(M)]
    inf = [i for i in range(M) if ref[i].in_frame]
    assert [got[i].cell for i in inf] == [ref[i].cell for i in inf]
    assert gsum.n_in_frame == rsum.n_in_frame and gsum.used_cell_all == rsum.used_cell_all
    same = all(got[i].tried == ref[i].tried and got[i].matched == ref[i].matched for i in range(M))
    if same:
        assert (gsum.n_matches, gsum.n_trials) == (rsum.n_matches, rsum.n_trials)
        dpx = [np.hypot(got[i].px[0] - ref[i].px[0], got[i].px[1] - ref[i].px[1]) for i in range(M) if ref[i].matched]
        assert all(got[i].order == ref[i].order and got[i].search_level == ref[i].search_level for i in range(M) if ref[i].matched)
        assert np.median(dpx) < 2e-3 and np.max(dpx) < 0.25
    else:
        diff = [i for i in range(M) if got[i].tried != ref[i].tried or got[i].matched != ref[i].matched]
        assert len(diff) <= 0.02 * max(rsum.n_trials, 1) + 2 and abs(gsum.n_matches - rsum.n_matches) <= 2, (len(diff), gsum.n_matches, rsum.n_matches)
    for k in kfs:
        k.close()
    cur.close()
    ctx.close()
