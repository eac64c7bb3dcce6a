"""ctypes binding of oracle/_ref/libhso_ref.so — the REFERENCE'S OWN hot-path sources (src/CoarseTracker.cpp, feature_alignment.cpp, matcher.cpp,
pose_optimizer.cpp, frame.cpp, point.cpp, camera.cpp, vikit/*, vendored Sophus) compiled unmodified against oracle/shim (see oracle/Makefile,
oracle/ref_wrap.cpp). Test infrastructure only: it pins the oracle restatement and serves bench.py's reference arm. The library is built in the
build container (where /root/reference exists) and travels to the GPU box as a prebuilt file; available() tells whether it is there."""
import ctypes as C
import os

import numpy as np

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libhso_ref.so")

_lib = None


def available():
    return os.path.exists(LIB)


def load():
    global _lib
    if _lib is None:
        O.load()  # liboracle_hso.so carries the cv2-pinned OpenCV restatements the shim forwards to
        lib = C.CDLL(LIB)
        lib.ref_describe.restype = C.c_char_p
        lib.ref_frame_new.restype = C.c_void_p
        lib.ref_coarse_track.restype = C.c_uint64
        lib.ref_mad_scale.restype = C.c_float
        lib.ref_huber_weight.restype = C.c_float
        lib.ref_get_median_f.restype = C.c_float
        lib.ref_get_median_d.restype = C.c_double
        lib.ref_error_multiplier2.restype = C.c_double
        for n in ("ref_align2d", "ref_align1d", "ref_get_best_search_level", "ref_check_ncc", "ref_create_pyramid", "ref_frame_n_levels"):
            getattr(lib, n).restype = C.c_int
        _lib = lib
    return _lib


dp = O.dp


def _rt(T):
    return np.ascontiguousarray(np.asarray(T, np.float64)[:3], np.float64).reshape(12).copy()


class Frame:
    """hso::Frame built by the reference's own constructor: pyramid (halfSample / cv::resize), Sobel images, statistics."""

    def __init__(self, cam, img, T_f_w=None, exposure_time=1.0, keyframe_id=0):
        lib = load()
        self.cam = O.cam_of(cam) if isinstance(cam, dict) else cam
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        T = _rt(np.eye(4) if T_f_w is None else T_f_w)
        self.h = lib.ref_frame_new(C.byref(self.cam), img.ctypes.data_as(C.c_void_p), W, H, dp(T), C.c_double(exposure_time), int(keyframe_id))
        if not self.h:
            raise ValueError("hso::Frame constructor threw (image does not match the camera model)")
        self.h = C.c_void_p(self.h)

    def close(self):
        if self.h:
            load().ref_frame_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        a, b = C.c_float(), C.c_float()
        load().ref_frame_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def levels(self):
        lib = load()
        out = []
        for l in range(lib.ref_frame_n_levels(self.h)):
            w, h = C.c_int(), C.c_int()
            lib.ref_frame_level_size(self.h, l, C.byref(w), C.byref(h))
            buf = np.zeros((h.value, w.value), np.uint8)
            lib.ref_frame_level(self.h, l, buf.ctypes.data_as(C.c_void_p))
            out.append(buf)
        return out

    def sobel(self, level):
        lib = load()
        w, h = C.c_int(), C.c_int()
        lib.ref_frame_level_size(self.h, level, C.byref(w), C.byref(h))
        gx = np.zeros((h.value, w.value), np.int16)
        gy = np.zeros((h.value, w.value), np.int16)
        lib.ref_frame_sobel(self.h, level, gx.ctypes.data_as(C.c_void_p), gy.ctypes.data_as(C.c_void_p))
        return gx, gy

    def set_pose(self, T):
        load().ref_frame_set_pose(self.h, dp(_rt(T)))

    def set_track_features(self, px, f, dist):
        px = np.ascontiguousarray(px, np.float64).reshape(-1)
        f = np.ascontiguousarray(f, np.float64).reshape(-1)
        dist = np.ascontiguousarray(dist, np.float64).reshape(-1)
        load().ref_frame_set_track_features(self.h, len(dist), dp(px), dp(f), dp(dist))


def coarse_track(ref, cur, T0, inverse_comp=False, max_level=4, min_level=1, n_iter=50):
    """CoarseTracker::run(ref, cur) of the reference. The initial exposure ratio is formed by run() itself from the frames' statistics."""
    lib = load()
    prm = O.orc_track_params(int(inverse_comp), max_level, min_level, n_iter)
    T = _rt(T0)
    a, et = C.c_float(), C.c_double()
    n = lib.ref_coarse_track(ref.h, cur.h, C.byref(prm), dp(T), C.byref(a), C.byref(et))
    return dict(T_cur_ref=T.reshape(3, 4), exposure_rat=a.value, n_tracked=int(n), exposure_time=et.value)


def track_eval(ref, cur, level, max_level, T, a, huber, outlier, inverse_comp=False):
    lib = load()
    H, b = np.zeros(49), np.zeros(7)
    E, tt, st = C.c_double(), C.c_int(), C.c_int()
    lib.ref_track_eval(ref.h, cur.h, int(inverse_comp), level, max_level, dp(_rt(T)), C.c_float(a), C.c_float(huber), C.c_float(outlier), dp(H), dp(b),
                       C.byref(E), C.byref(tt), C.byref(st))
    return H.reshape(7, 7), b, E.value, tt.value, st.value


def track_select_robust(ref, cur, level, max_level, T, a):
    hu, ou = C.c_float(), C.c_float()
    load().ref_track_select_robust(ref.h, cur.h, level, max_level, dp(_rt(T)), C.c_float(a), C.byref(hu), C.byref(ou))
    return hu.value, ou.value


def make_depth_ref(ref, hosts, host_of, has_point, f_host, idist):
    lib = load()
    F = len(idist)
    hh = (C.c_void_p * F)(*[(hosts[host_of[i]].h if has_point[i] else None) for i in range(F)])
    out = np.zeros(F)
    lib.ref_make_depth_ref(ref.h, F, hh, dp(np.ascontiguousarray(f_host, np.float64)), dp(np.ascontiguousarray(idist, np.float64)), dp(out))
    return out


def half_sample(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h // 2, w // 2), np.uint8)
    load().ref_half_sample(img.ctypes.data_as(C.c_void_p), w, h, out.ctypes.data_as(C.c_void_p))
    return out


def create_pyramid(img, n_levels=5):
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    out = np.zeros(2 * W * H, np.uint8)
    lw, lh = (C.c_int * n_levels)(), (C.c_int * n_levels)()
    load().ref_create_pyramid(img.ctypes.data_as(C.c_void_p), W, H, n_levels, out.ctypes.data_as(C.c_void_p), lw, lh)
    levels, o = [], 0
    for l in range(n_levels):
        n = lw[l] * lh[l]
        levels.append(out[o:o + n].reshape(lh[l], lw[l]).copy())
        o += n
    return levels


def accumulator7(J, w):
    J = np.ascontiguousarray(J, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    H = np.zeros(49, np.float32)
    load().ref_accumulator7(len(w), J.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), H.ctypes.data_as(C.c_void_p))
    return H.reshape(7, 7)


def mad_scale(errors):
    e = np.ascontiguousarray(errors, np.float32)
    return load().ref_mad_scale(e.ctypes.data_as(C.c_void_p), len(e))


def huber_weight(x):
    return load().ref_huber_weight(C.c_float(x))


def get_median(v):
    v = np.ascontiguousarray(v)
    if v.dtype == np.float32:
        return load().ref_get_median_f(v.ctypes.data_as(C.c_void_p), len(v))
    v = v.astype(np.float64)
    return load().ref_get_median_d(v.ctypes.data_as(C.c_void_p), len(v))


def world2cam(cam, xyz):
    c = O.cam_of(cam)
    out = np.zeros(2)
    load().ref_world2cam(C.byref(c), dp(np.ascontiguousarray(xyz, np.float64)), dp(out))
    return out


def cam2world(cam, u, v):
    c = O.cam_of(cam)
    out = np.zeros(3)
    load().ref_cam2world(C.byref(c), C.c_double(u), C.c_double(v), dp(out))
    return out


def se3_exp(t):
    out = np.zeros(12)
    load().ref_se3_exp(dp(np.ascontiguousarray(t, np.float64)), dp(out))
    return out.reshape(3, 4)


def se3_log(T):
    out = np.zeros(6)
    load().ref_se3_log(dp(_rt(T)), dp(out))
    return out


def se3_mul(A, B):
    out = np.zeros(12)
    load().ref_se3_mul(dp(_rt(A)), dp(_rt(B)), dp(out))
    return out.reshape(3, 4)


def se3_inverse(A):
    out = np.zeros(12)
    load().ref_se3_inverse(dp(_rt(A)), dp(out))
    return out.reshape(3, 4)


def align2d(img, ref_patch_with_border, ref_patch, px, n_iter=10):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    rb = np.ascontiguousarray(ref_patch_with_border, np.float32)
    rp = np.ascontiguousarray(ref_patch, np.float32)
    p = np.array(px, np.float64)
    cur = np.zeros(64, np.float32)
    ok = load().ref_align2d(img.ctypes.data_as(C.c_void_p), w, h, w, rb.ctypes.data_as(C.c_void_p), rp.ctypes.data_as(C.c_void_p), n_iter, dp(p),
                            cur.ctypes.data_as(C.c_void_p))
    return bool(ok), p, cur


def align1d(img, direction, ref_patch_with_border, ref_patch, px, n_iter=10):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    d = np.ascontiguousarray(direction, np.float32)
    rb = np.ascontiguousarray(ref_patch_with_border, np.float32)
    rp = np.ascontiguousarray(ref_patch, np.float32)
    p = np.array(px, np.float64)
    cur = np.zeros(64, np.float32)
    hinv = C.c_double()
    ok = load().ref_align1d(img.ctypes.data_as(C.c_void_p), w, h, w, d.ctypes.data_as(C.c_void_p), rb.ctypes.data_as(C.c_void_p),
                            rp.ctypes.data_as(C.c_void_p), n_iter, dp(p), C.byref(hinv), cur.ctypes.data_as(C.c_void_p))
    return bool(ok), p, hinv.value, cur


def get_warp_matrix_affine(cam, px_ref, f_ref, depth_ref, T_cur_ref, level_ref):
    c = O.cam_of(cam)
    A = np.zeros(4)
    load().ref_get_warp_matrix_affine(C.byref(c), dp(np.ascontiguousarray(px_ref, np.float64)), dp(np.ascontiguousarray(f_ref, np.float64)),
                                      C.c_double(depth_ref), dp(_rt(T_cur_ref)), int(level_ref), dp(A))
    return A.reshape(2, 2)


def warp_affine(A, img_ref, px_ref, level_ref, search_level, halfpatch=5):
    img = np.ascontiguousarray(img_ref, np.uint8)
    h, w = img.shape
    out = np.zeros((2 * halfpatch) ** 2, np.float32)
    load().ref_warp_affine(dp(np.ascontiguousarray(A, np.float64).reshape(4)), img.ctypes.data_as(C.c_void_p), w, h,
                           dp(np.ascontiguousarray(px_ref, np.float64)), level_ref, search_level, halfpatch, out.ctypes.data_as(C.c_void_p))
    return out


def find_match_batch(cur, kfs, cands, px_init, seed_mode=False):
    """Matcher::findMatchDirect (or findMatchSeed) for every candidate record (orc_reproj_cand array). Returns ok, px, search_level, A, h_inv."""
    lib = load()
    M = len(px_init)
    px = np.ascontiguousarray(px_init, np.float64).reshape(-1).copy()
    ok = np.zeros(M, np.int32)
    sl = np.zeros(M, np.int32)
    A = np.zeros(4 * M)
    hinv = np.zeros(M)
    hh = (C.c_void_p * len(kfs))(*[k.h for k in kfs])
    lib.ref_find_match_batch(cur.h, len(kfs), hh, M, cands, int(seed_mode), dp(px), ok.ctypes.data_as(C.c_void_p), sl.ctypes.data_as(C.c_void_p), dp(A), dp(hinv))
    return ok, px.reshape(M, 2), sl, A.reshape(M, 2, 2), hinv


def find_match_seed_batch(cur, kfs, seeds, px_init):
    """Matcher::findMatchSeed for every seed record (orc_seed_obs array). Returns ok, px, search_level, A."""
    lib = load()
    S = len(px_init)
    px = np.ascontiguousarray(px_init, np.float64).reshape(-1).copy()
    ok = np.zeros(S, np.int32)
    sl = np.zeros(S, np.int32)
    A = np.zeros(4 * S)
    hh = (C.c_void_p * len(kfs))(*[k.h for k in kfs])
    lib.ref_find_match_seed_batch(cur.h, len(kfs), hh, S, seeds, dp(px), ok.ctypes.data_as(C.c_void_p), sl.ctypes.data_as(C.c_void_p), dp(A))
    return ok, px.reshape(S, 2), sl, A.reshape(S, 2, 2)


def shi_tomasi(img, xy):
    """hso::shiTomasiScore (src/vikit/vision.cpp:111-151) at the pixels xy (n,2)."""
    lib = load()
    img = np.ascontiguousarray(img, np.uint8)
    xy = np.ascontiguousarray(xy, np.int32)
    out = np.zeros(len(xy), np.float32)
    lib.ref_shi_tomasi(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], len(xy), xy.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def reproject_match(cur, kfs, cands, grid, cell_order):
    """The grid stage of Reprojector::reprojectMap through the reference's own reprojectPoint / reprojectCell / reprojectCellAll
    (oracle/ref_wrap.cpp: ref_reproject_match). cands: orc_reproj_cand array, grid: orc_reproj_grid. Returns (orc_reproj_result array, summary)."""
    import oracle_lib as O
    lib = load()
    M = len(cands)
    out = (O.orc_reproj_result * max(M, 1))()
    summ = O.orc_reproj_summary()
    order = np.ascontiguousarray(cell_order, np.int32)
    hh = (C.c_void_p * len(kfs))(*[k.h for k in kfs])
    lib.ref_reproject_match(cur.h, len(kfs), hh, M, cands, C.byref(grid), order.ctypes.data_as(C.c_void_p), out, C.byref(summ))
    return out, summ


def reproject_seeds(cur, kfs, seeds, grid, cell_order, n_matches_in):
    """The seed stage of Reprojector::reprojectMap through the reference's own reprojectorSeed / reprojectorSeeds (ref_reproject_seeds).
    Returns (orc_reproj_result array — matched / order / px / search level; tried only where matched —, summary with n_in_frame, n_matches)."""
    import oracle_lib as O
    lib = load()
    S = len(seeds)
    out = (O.orc_reproj_result * max(S, 1))()
    summ = O.orc_reproj_summary()
    order = np.ascontiguousarray(cell_order, np.int32)
    hh = (C.c_void_p * len(kfs))(*[k.h for k in kfs])
    lib.ref_reproject_seeds(cur.h, len(kfs), hh, S, seeds, C.byref(grid), order.ctypes.data_as(C.c_void_p), int(n_matches_in), out, C.byref(summ))
    return out, summ


def depth_observe(cur, kfs, seeds, px_error_angle):
    """DepthFilter::observeDepthRow (src/depth_filter.cpp:580-675) of the reference for every seed record (orc_seed_obs array), one seed per call.
    Returns an orc_seed_result array (z = 1 / mu after the update, as the reference records it in Seed::vec_distance)."""
    import oracle_lib as O
    lib = load()
    S = len(seeds)
    out = (O.orc_seed_result * max(S, 1))()
    hh = (C.c_void_p * len(kfs))(*[k.h for k in kfs])
    lib.ref_depth_observe(cur.h, len(kfs), hh, C.c_double(px_error_angle), S, seeds, out)
    return out


def pose_optimize(cam, p, reproj_thresh=2.0, n_iter=12, blank=None):
    """optimizeLevenbergMarquardt3rd of the reference on a synth.make_pose_problem record."""
    lib = load()
    W, H = cam["width"], cam["height"]
    img = np.zeros((H, W), np.uint8) if blank is None else blank
    fr = Frame(cam, img)
    K = len(p["T_host_w"])
    hosts = [Frame(cam, img) for _ in range(K)]
    F = len(p["level"])
    f = np.ascontiguousarray(p["f"], np.float64).reshape(-1)
    ph = np.ascontiguousarray(p["p_host"], np.float64).reshape(-1)
    hi = np.ascontiguousarray(p["host_idx"], np.int32)
    Th = np.ascontiguousarray(np.asarray(p["T_host_w"], np.float64)[:, :3], np.float64).reshape(-1)
    g = np.ascontiguousarray(p["grad"], np.float64).reshape(-1)
    lv = np.ascontiguousarray(p["level"], np.int8)
    ft = np.ascontiguousarray(p["ftype"], np.int8)
    pt = np.ascontiguousarray(p["ptype"], np.int8)
    outl = np.zeros(max(F, 1), np.uint8)
    out = O.orc_pose_result()
    hh = (C.c_void_p * K)(*[h.h for h in hosts])
    lib.ref_pose_optimize(fr.h, K, hh, C.c_double(reproj_thresh), n_iter, int(p.get("n_fts_total", F)), F, dp(f), dp(ph), hi.ctypes.data_as(C.c_void_p), dp(Th),
                          dp(g), lv.ctypes.data_as(C.c_void_p), ft.ctypes.data_as(C.c_void_p), pt.ctypes.data_as(C.c_void_p), dp(_rt(p["T_f_w"])),
                          outl.ctypes.data_as(C.c_void_p), C.byref(out))
    res = dict(T_f_w=np.array(out.T_f_w[:]).reshape(3, 4), cov=np.array(out.cov[:]).reshape(6, 6), estimated_scale=out.estimated_scale,
               error_init=out.error_init, error_final=out.error_final, num_obs=int(out.num_obs), error_in_px=out.error_in_px, outlier=outl[:F].copy())
    fr.close()
    for h in hosts:
        h.close()
    return res
