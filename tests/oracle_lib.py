"""ctypes binding of the CPU oracle (oracle/liboracle_hso.so) — test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle_hso.so")


class orc_cam(C.Structure):
    _fields_ = [("model", C.c_int), ("width", C.c_int), ("height", C.c_int), ("undistort", C.c_int),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("d", C.c_double * 5)]


class orc_trace(C.Structure):
    _fields_ = [("level", C.c_int), ("iter", C.c_int), ("T_eval", C.c_double * 12), ("a_eval", C.c_float), ("lambda_", C.c_float),
                ("H", C.c_double * 49), ("b", C.c_double * 7), ("step", C.c_double * 7), ("energy", C.c_double),
                ("total_terms", C.c_int), ("saturated_terms", C.c_int), ("accepted", C.c_int), ("huber", C.c_float), ("outlier", C.c_float)]


class orc_track_params(C.Structure):
    _fields_ = [("inverse_comp", C.c_int), ("max_level", C.c_int), ("min_level", C.c_int), ("n_iter", C.c_int)]


class orc_align_job(C.Structure):
    _fields_ = [("ref_level", C.c_int32), ("search_level", C.c_int32), ("type", C.c_int32), ("scale_patch", C.c_int32),
                ("px_ref", C.c_double * 2), ("A_cur_ref", C.c_double * 4), ("grad", C.c_double * 2), ("px_cur", C.c_double * 2),
                ("exposure_rat", C.c_float), ("ncc_thresh", C.c_float)]


class orc_align_result(C.Structure):
    _fields_ = [("ok", C.c_int32), ("align_converged", C.c_int32), ("px_cur", C.c_double * 2), ("h_inv", C.c_double)]


class orc_pose_result(C.Structure):
    _fields_ = [("T_f_w", C.c_double * 12), ("cov", C.c_double * 36), ("estimated_scale", C.c_double), ("error_init", C.c_double),
                ("error_final", C.c_double), ("num_obs", C.c_uint64), ("error_in_px", C.c_float), ("n_trials_total", C.c_int),
                ("early_return", C.c_int)]


_lib = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_coarse_track.restype = C.c_uint64
        for n in ("orc_align2d", "orc_align1d", "orc_get_best_search_level", "orc_check_ncc", "orc_check_normal", "orc_create_pyramid"):
            getattr(_lib, n).restype = C.c_int
    return _lib


def cam_of(c):
    o = orc_cam()
    o.model, o.width, o.height, o.undistort = c.get("model", 0), c["width"], c["height"], c.get("undistort", 0)
    o.fx, o.fy, o.cx, o.cy = c["fx"], c["fy"], c["cx"], c["cy"]
    for i in range(5):
        o.d[i] = float(c["d"][i])
    return o


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pad_level(img, pad_rows=3):
    """Level image followed by zero rows: the reference's forward-mode gradient taps may read row == rows (undefined there);
    both the oracle harness and the CUDA path define those bytes as zero."""
    h, w = img.shape
    buf = np.zeros((h + pad_rows) * w + 64, np.uint8)
    buf[: h * w] = img.reshape(-1)
    return buf


def create_pyramid(img, n_levels=5):
    lib = load()
    H, W = img.shape
    out = np.zeros(W * H, np.uint8)
    lw = (C.c_int * n_levels)()
    lh = (C.c_int * n_levels)()
    img = np.ascontiguousarray(img)
    path = lib.orc_create_pyramid(img.ctypes.data_as(C.c_void_p), W, H, n_levels, out.ctypes.data_as(C.c_void_p), lw, lh)
    levels = [img]
    o = 0
    for i in range(1, n_levels):
        n = lw[i] * lh[i]
        levels.append(out[o:o + n].reshape(lh[i], lw[i]).copy())
        o += n
    return levels, path


def sobel5(img):
    lib = load()
    h, w = img.shape
    gx = np.zeros((h, w), np.int16)
    gy = np.zeros((h, w), np.int16)
    img = np.ascontiguousarray(img)
    lib.orc_sobel5(img.ctypes.data_as(C.c_void_p), w, h, gx.ctypes.data_as(C.c_void_p), gy.ctypes.data_as(C.c_void_p))
    return gx, gy


def frame_stats(img):
    lib = load()
    gx, gy = sobel5(img)
    h, w = img.shape
    a, b = C.c_float(), C.c_float()
    img = np.ascontiguousarray(img)
    lib.orc_frame_stats(img.ctypes.data_as(C.c_void_p), gx.ctypes.data_as(C.c_void_p), gy.ctypes.data_as(C.c_void_p), w, h, C.byref(a), C.byref(b))
    return a.value, b.value


class TrackProblem:
    """Flattened CoarseTracker inputs for the oracle: padded level buffers + feature arrays."""

    def __init__(self, cam, ref_levels, cur_levels, px, f, dist):
        self.cam = cam_of(cam)
        self.n = len(ref_levels)
        self.ref_bufs = [pad_level(l) for l in ref_levels]
        self.cur_bufs = [pad_level(l) for l in cur_levels]
        self.refp = (C.c_void_p * self.n)(*[b.ctypes.data for b in self.ref_bufs])
        self.curp = (C.c_void_p * self.n)(*[b.ctypes.data for b in self.cur_bufs])
        self.lw = (C.c_int * self.n)(*[l.shape[1] for l in ref_levels])
        self.lh = (C.c_int * self.n)(*[l.shape[0] for l in ref_levels])
        self.px = np.ascontiguousarray(px, np.float64).reshape(-1)
        self.f = np.ascontiguousarray(f, np.float64).reshape(-1)
        self.dist = np.ascontiguousarray(dist, np.float64).reshape(-1)
        self.F = self.dist.shape[0]

    def run(self, T0, a0, inverse_comp=False, max_level=4, min_level=1, n_iter=50, trace_cap=512):
        lib = load()
        prm = orc_track_params(int(inverse_comp), max_level, min_level, n_iter)
        T = np.ascontiguousarray(T0, np.float64).reshape(12).copy()
        a = C.c_float(a0)
        trace = (orc_trace * trace_cap)()
        tl, ne = C.c_int(), C.c_int()
        n = lib.orc_coarse_track(C.byref(self.cam), C.byref(prm), self.n, self.refp, self.curp, self.lw, self.lh, self.F, dp(self.px),
                                 dp(self.f), dp(self.dist), dp(T), C.byref(a), trace, trace_cap, C.byref(tl), C.byref(ne))
        return dict(T_cur_ref=T.reshape(3, 4), exposure_rat=a.value, n_tracked=int(n), n_evals=ne.value, trace=[trace[i] for i in range(tl.value)])

    def eval(self, level, max_level, T, a, huber, outlier, inverse_comp=False):
        lib = load()
        H = np.zeros(49)
        b = np.zeros(7)
        E = C.c_double()
        tt, st = C.c_int(), C.c_int()
        T = np.ascontiguousarray(T, np.float64).reshape(12)
        lib.orc_track_eval(C.byref(self.cam), int(inverse_comp), level, max_level, self.refp[level], self.curp[level], self.lw[level],
                           self.lh[level], self.F, dp(self.px), dp(self.f), dp(self.dist), dp(T), C.c_float(a), C.c_float(huber),
                           C.c_float(outlier), dp(H), dp(b), C.byref(E), C.byref(tt), C.byref(st))
        return H.reshape(7, 7), b, E.value, tt.value, st.value

    def select_robust(self, level, max_level, T, a):
        lib = load()
        hu, ou = C.c_float(), C.c_float()
        n = C.c_int()
        T = np.ascontiguousarray(T, np.float64).reshape(12)
        lib.orc_track_select_robust(C.byref(self.cam), level, max_level, self.refp[level], self.curp[level], self.lw[level], self.lh[level],
                                    self.F, dp(self.px), dp(self.f), dp(self.dist), dp(T), C.c_float(a), C.byref(hu), C.byref(ou), C.byref(n))
        return hu.value, ou.value, n.value


def make_depth_ref(T_ref_w, has_point, f_host, idist, T_host_w):
    """CoarseTracker::makeDepthRef (src/CoarseTracker.cpp:210-240): T_host_w is one 3x4 pose PER FEATURE (its point's host frame)."""
    lib = load()
    F = len(idist)
    out = np.zeros(F)
    hp = np.ascontiguousarray(has_point, np.uint8)
    lib.orc_make_depth_ref(dp(np.ascontiguousarray(T_ref_w, np.float64).reshape(12)), F, hp.ctypes.data_as(C.POINTER(C.c_uint8)),
                           dp(np.ascontiguousarray(f_host, np.float64)), dp(np.ascontiguousarray(idist, np.float64)),
                           dp(np.ascontiguousarray(T_host_w, np.float64)), dp(out))
    return out


def track_solve(H, b, lam):
    lib = load()
    H = np.ascontiguousarray(H, np.float64).reshape(49)
    b = np.ascontiguousarray(b, np.float64).reshape(7)
    step = np.zeros(7)
    lib.orc_track_solve(dp(H), dp(b), C.c_float(lam), dp(step))
    return step


def match_direct_batch(jobs, ref_levels, cur_levels, cur_sobel, align_max_iter=10):
    """jobs: list of dicts (hso_b200.synth.make_align_jobs). ref/cur_levels: lists of u8 level images; cur_sobel: [(gx, gy)] x3."""
    lib = load()
    M = len(jobs)
    arr = (orc_align_job * M)()
    for m, j in enumerate(jobs):
        a = arr[m]
        a.ref_level, a.search_level, a.type, a.scale_patch = j["ref_level"], j["search_level"], j["type"], j.get("scale_patch", 0)
        for k in range(2):
            a.px_ref[k], a.grad[k], a.px_cur[k] = j["px_ref"][k], j["grad"][k], j["px_cur"][k]
        A = np.asarray(j["A_cur_ref"], np.float64).reshape(4)
        for k in range(4):
            a.A_cur_ref[k] = A[k]
        a.exposure_rat = j.get("exposure_rat", 1.0)
    n = len(ref_levels)
    rb = [pad_level(l) for l in ref_levels]
    cb = [pad_level(l) for l in cur_levels]
    refp = (C.c_void_p * n)(*[b.ctypes.data for b in rb])
    curp = (C.c_void_p * n)(*[b.ctypes.data for b in cb])
    lw = (C.c_int * n)(*[l.shape[1] for l in ref_levels])
    lh = (C.c_int * n)(*[l.shape[0] for l in ref_levels])
    sx = [np.ascontiguousarray(g[0]) for g in cur_sobel]
    sy = [np.ascontiguousarray(g[1]) for g in cur_sobel]
    sxp = (C.c_void_p * 3)(*[a.ctypes.data for a in sx])
    syp = (C.c_void_p * 3)(*[a.ctypes.data for a in sy])
    out = (orc_align_result * M)()
    lib.orc_match_direct_batch(M, arr, refp, curp, lw, lh, sxp, syp, align_max_iter, out)
    return out


def pose_optimize(p, reproj_thresh=2.0, n_iter=12):
    lib = load()
    f = np.ascontiguousarray(p["f"], np.float64).reshape(-1)
    ph = np.ascontiguousarray(p["p_host"], np.float64).reshape(-1)
    hi = np.ascontiguousarray(p["host_idx"], np.int32)
    Th = np.ascontiguousarray(p["T_host_w"], np.float64).reshape(-1)
    g = np.ascontiguousarray(p["grad"], np.float64).reshape(-1)
    lv, ft, pt = (np.ascontiguousarray(p[k], np.int8) for k in ("level", "ftype", "ptype"))
    T0 = np.ascontiguousarray(p["T_f_w"], np.float64).reshape(12)
    F = hi.shape[0]
    outl = np.zeros(max(F, 1), np.uint8)
    out = orc_pose_result()
    i8 = C.POINTER(C.c_int8)
    lib.orc_pose_optimize(C.c_double(reproj_thresh), n_iter, C.c_double(p["err_mult2"]), int(p["n_fts_total"]), F, dp(f), dp(ph),
                          hi.ctypes.data_as(C.POINTER(C.c_int32)), dp(Th), dp(g), lv.ctypes.data_as(i8), ft.ctypes.data_as(i8),
                          pt.ctypes.data_as(i8), dp(T0), outl.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(out))
    return dict(T_f_w=np.array(out.T_f_w[:]).reshape(3, 4), cov=np.array(out.cov[:]).reshape(6, 6), estimated_scale=out.estimated_scale,
                error_init=out.error_init, error_final=out.error_final, num_obs=int(out.num_obs), error_in_px=float(out.error_in_px),
                n_trials_total=out.n_trials_total, early_return=out.early_return, outlier=outl[:F].copy())


# ---- N2: FAST-9 (real reference library when oracle/_ref/libfast_ref.so exists, restatement always) -----------------------------------
class orc_corner(C.Structure):
    _fields_ = [("x", C.c_int16), ("y", C.c_int16), ("score", C.c_int32), ("shi_tomasi", C.c_float)]


REF_FAST_LIB = os.path.join(ORACLE_DIR, "_ref", "libfast_ref.so")
_ref_fast = None


def ref_fast_available():
    return os.path.exists(REF_FAST_LIB)


def ref_fast9(img, threshold):
    """The REAL reference: fast_corner_detect_9_sse2 + fast_corner_score_9 + fast_nonmax_3x3. Returns (xy (n,2), scores (n,), nonmax idx)."""
    global _ref_fast
    if _ref_fast is None:
        _ref_fast = C.CDLL(REF_FAST_LIB)
        _ref_fast.ref_fast9_detect.restype = C.c_int
    img = np.ascontiguousarray(img)
    h, w = img.shape
    cap = w * h
    xy = np.zeros((cap, 2), np.int16)
    sc = np.zeros(cap, np.int32)
    nm = np.zeros(cap, np.int32)
    nnm = C.c_int()
    n = _ref_fast.ref_fast9_detect(img.ctypes.data_as(C.c_void_p), w, h, w, int(threshold), xy.ctypes.data_as(C.c_void_p),
                                   sc.ctypes.data_as(C.c_void_p), nm.ctypes.data_as(C.c_void_p), cap, C.byref(nnm))
    return xy[:n].copy(), sc[:n].copy(), nm[:nnm.value].copy()


def fast9_corners(img, threshold):
    lib = load()
    lib.orc_fast9_corners.restype = C.c_int
    img = np.ascontiguousarray(img)
    h, w = img.shape
    cap = w * h
    xy = np.zeros((cap, 2), np.int16)
    sc = np.zeros(cap, np.int32)
    n = lib.orc_fast9_corners(img.ctypes.data_as(C.c_void_p), w, h, w, int(threshold), xy.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), cap)
    return xy[:n].copy(), sc[:n].copy()


def fast_detect(img, threshold, border=8):
    lib = load()
    lib.orc_fast_detect.restype = C.c_int
    img = np.ascontiguousarray(img)
    h, w = img.shape
    cap = w * h // 4 + 16
    out = (orc_corner * cap)()
    n = lib.orc_fast_detect(img.ctypes.data_as(C.c_void_p), w, h, w, int(threshold), int(border), out, cap)
    return [(out[i].x, out[i].y, out[i].score, out[i].shi_tomasi) for i in range(n)]


# ---- N1: Reprojector::reprojectMap data path + Matcher::findMatchDirect ------------------------------------------------------------
class orc_reproj_cand(C.Structure):
    _fields_ = [("p_host", C.c_double * 3), ("px_ref", C.c_double * 2), ("f_ref", C.c_double * 3), ("grad", C.c_double * 2),
                ("depth_ref", C.c_double), ("host_pose", C.c_int32), ("ref_pose", C.c_int32), ("ref_frame", C.c_int32),
                ("ref_level", C.c_int32), ("ftr_type", C.c_int32), ("pt_type", C.c_int32), ("pt_ftr_type", C.c_int32),
                ("scale_patch", C.c_int32), ("exposure_rat", C.c_float), ("pad_", C.c_float)]


class orc_reproj_grid(C.Structure):
    _fields_ = [("cell_size", C.c_int32), ("n_cols", C.c_int32), ("n_rows", C.c_int32), ("max_fts", C.c_int32),
                ("align_max_iter", C.c_int32), ("pad_", C.c_int32)]


class orc_reproj_result(C.Structure):
    _fields_ = [("in_frame", C.c_int32), ("cell", C.c_int32), ("tried", C.c_int32), ("matched", C.c_int32), ("search_level", C.c_int32),
                ("order", C.c_int32), ("align_ok", C.c_int32), ("pad_", C.c_int32), ("px", C.c_double * 2), ("A_cur_ref", C.c_double * 4)]


class orc_reproj_summary(C.Structure):
    _fields_ = [("n_in_frame", C.c_int32), ("n_matches", C.c_int32), ("n_trials", C.c_int32), ("used_cell_all", C.c_int32)]


def cam2world(cam, u, v):
    lib = load()
    out = np.zeros(3)
    lib.orc_cam2world(C.byref(cam_of(cam)), C.c_double(u), C.c_double(v), dp(out))
    return out


def get_warp_matrix_affine(cam, px_ref, f_ref, depth_ref, T_cur_ref, level_ref):
    lib = load()
    A = np.zeros(4)
    px = np.ascontiguousarray(px_ref, np.float64)
    f = np.ascontiguousarray(f_ref, np.float64)
    T = np.ascontiguousarray(T_cur_ref, np.float64).reshape(12)
    lib.orc_get_warp_matrix_affine(C.byref(cam_of(cam)), dp(px), dp(f), C.c_double(depth_ref), dp(T), int(level_ref), dp(A))
    return A.reshape(2, 2)


def _reproj_args(T_cur_w, T_f_w, ref_pyramids, cur_levels, cur_sobel):
    T = np.ascontiguousarray(T_cur_w, np.float64).reshape(12)
    Tk = np.ascontiguousarray(T_f_w, np.float64).reshape(-1)
    nl = len(cur_levels)
    keep = [T, Tk]
    frames = (C.POINTER(C.c_void_p) * len(ref_pyramids))()
    for i, pyr in enumerate(ref_pyramids):
        lv = [np.ascontiguousarray(l) for l in pyr]
        arr = (C.c_void_p * nl)(*[a.ctypes.data for a in lv])
        keep += [lv, arr]
        frames[i] = C.cast(arr, C.POINTER(C.c_void_p))
    cl = [np.ascontiguousarray(l) for l in cur_levels]
    curp = (C.c_void_p * nl)(*[a.ctypes.data for a in cl])
    lw = (C.c_int * nl)(*[l.shape[1] for l in cl])
    lh = (C.c_int * nl)(*[l.shape[0] for l in cl])
    sx = [np.ascontiguousarray(g[0]) for g in cur_sobel]
    sy = [np.ascontiguousarray(g[1]) for g in cur_sobel]
    sxp = (C.c_void_p * 3)(*[a.ctypes.data for a in sx])
    syp = (C.c_void_p * 3)(*[a.ctypes.data for a in sy])
    keep += [cl, sx, sy]
    return T, Tk, frames, curp, lw, lh, sxp, syp, keep


def reproject_match(cam, T_cur_w, T_f_w, cands, grid, cell_order, max_search_level, ref_pyramids, cur_levels, cur_sobel):
    """cands: ctypes array of orc_reproj_cand (same layout as hso_reproj_cand); ref_pyramids: list (by ref_frame index) of lists of level
    images; cur_sobel: [(gx, gy)] for levels 0..2. Returns (results array, summary)."""
    lib = load()
    M = len(cands)
    T, Tk, frames, curp, lw, lh, sxp, syp, keep = _reproj_args(T_cur_w, T_f_w, ref_pyramids, cur_levels, cur_sobel)
    order = np.ascontiguousarray(cell_order, np.int32)
    out = (orc_reproj_result * max(M, 1))()
    summ = orc_reproj_summary()
    lib.orc_reproject_match(C.byref(cam_of(cam)), dp(T), Tk.size // 12, dp(Tk), M, cands, C.byref(grid), order.ctypes.data_as(C.POINTER(C.c_int32)),
                            int(max_search_level), frames, curp, lw, lh, sxp, syp, out, C.byref(summ))
    return out, summ


def reproject_speculative(cam, T_cur_w, T_f_w, cands, grid, max_search_level, ref_pyramids, cur_levels, cur_sobel):
    """findMatchDirect for every candidate that entered a cell. Returns (results array with align_ok / A / search_level, px_after (M,2))."""
    lib = load()
    M = len(cands)
    T, Tk, frames, curp, lw, lh, sxp, syp, keep = _reproj_args(T_cur_w, T_f_w, ref_pyramids, cur_levels, cur_sobel)
    out = (orc_reproj_result * max(M, 1))()
    px_after = np.zeros((max(M, 1), 2))
    lib.orc_reproject_speculative(C.byref(cam_of(cam)), dp(T), Tk.size // 12, dp(Tk), M, cands, C.byref(grid), int(max_search_level), frames, curp,
                                  lw, lh, sxp, syp, out, dp(px_after))
    return out, px_after


def reproject_select(cands, match_ok, grid, cell_order, io):
    """The selection walk with findMatchDirect's outcome supplied (match_ok uint8 per candidate); io: results array whose in_frame / cell
    are inputs. Returns the summary; tried / matched / order are written into io."""
    lib = load()
    ok = np.ascontiguousarray(match_ok, np.uint8)
    order = np.ascontiguousarray(cell_order, np.int32)
    summ = orc_reproj_summary()
    lib.orc_reproject_select(len(cands), cands, ok.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(grid), order.ctypes.data_as(C.POINTER(C.c_int32)),
                             io, C.byref(summ))
    return summ


# ---- N4: undistortion maps + cv::remap -----------------------------------------------------------------------------------------------
def init_undistort_maps(cam):
    lib = load()
    W, H = cam["width"], cam["height"]
    m1 = np.zeros((H, W, 2), np.int16)
    m2 = np.zeros((H, W), np.uint16)
    lib.orc_init_undistort_maps.restype = C.c_int
    rc = lib.orc_init_undistort_maps(C.byref(cam_of(cam)), m1.ctypes.data_as(C.c_void_p), m2.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return m1, m2


def remap_linear(src, m1, m2):
    lib = load()
    src = np.ascontiguousarray(src, np.uint8)
    m1 = np.ascontiguousarray(m1, np.int16)
    m2 = np.ascontiguousarray(m2, np.uint16)
    dh, dw = m2.shape
    dst = np.zeros((dh, dw), np.uint8)
    lib.orc_remap_linear_u8(src.ctypes.data_as(C.c_void_p), src.shape[1], src.shape[0], src.shape[1], m1.ctypes.data_as(C.c_void_p),
                            m2.ctypes.data_as(C.c_void_p), dw, dh, dst.ctypes.data_as(C.c_void_p))
    return dst


def resize_linear(src, dw, dh):
    """cv::resize(INTER_LINEAR) restatement (orc_resize_linear_u8, pinned against cv2 golden vectors in test_oracle_pins.py)."""
    lib = load()
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib.orc_resize_linear_u8(src.ctypes.data_as(C.c_void_p), src.shape[1], src.shape[0], dst.ctypes.data_as(C.c_void_p), int(dw), int(dh))
    return dst


def convert_maps(mx, my):
    lib = load()
    mx = np.ascontiguousarray(mx, np.float32)
    my = np.ascontiguousarray(my, np.float32)
    h, w = mx.shape
    m1 = np.zeros((h, w, 2), np.int16)
    m2 = np.zeros((h, w), np.uint16)
    lib.orc_convert_maps(mx.ctypes.data_as(C.c_void_p), my.ctypes.data_as(C.c_void_p), h * w, m1.ctypes.data_as(C.c_void_p), m2.ctypes.data_as(C.c_void_p))
    return m1, m2


# ---- N3: DepthFilter::observeDepthRow --------------------------------------------------------------------------------------------------
class orc_seed_obs(C.Structure):
    _fields_ = [("px", C.c_double * 2), ("f", C.c_double * 3), ("grad", C.c_double * 2), ("ref_frame", C.c_int32), ("ref_pose", C.c_int32),
                ("level", C.c_int32), ("ftr_type", C.c_int32), ("mu", C.c_float), ("sigma2", C.c_float), ("exposure_rat", C.c_float),
                ("pad_", C.c_float)]


class orc_seed_result(C.Structure):
    _fields_ = [("is_update", C.c_int32), ("is_valid", C.c_int32), ("res", C.c_int32), ("search_level", C.c_int32), ("epl_start", C.c_int32 * 2),
                ("epl_end", C.c_int32 * 2), ("mu", C.c_float), ("sigma2", C.c_float), ("z", C.c_double), ("px_cur", C.c_double * 2)]


def depth_observe(cam, T_cur_w, T_f_w, seeds, px_error_angle, ref_pyramids, cur_levels, cur_sobel, max_search_level=2, align_max_iter=10):
    """seeds: ctypes array of orc_seed_obs (same layout as hso_seed_obs; ref_frame indexes ref_pyramids)."""
    lib = load()
    S = len(seeds)
    T, Tk, frames, curp, lw, lh, sxp, syp, keep = _reproj_args(T_cur_w, T_f_w, ref_pyramids, cur_levels, cur_sobel)
    out = (orc_seed_result * max(S, 1))()
    lib.orc_depth_observe(C.byref(cam_of(cam)), dp(T), Tk.size // 12, dp(Tk), C.c_double(px_error_angle), S, seeds, int(max_search_level),
                          int(align_max_iter), frames, curp, lw, lh, sxp, syp, out)
    return out


# ---- a13b: seed stage of Reprojector::reprojectMap ------------------------------------------------------------------------------------
def reproject_seeds(cam, T_cur_w, T_f_w, seeds, grid, cell_order, n_matches_in, max_search_level, ref_pyramids, cur_levels, cur_sobel):
    """seeds: ctypes array of orc_seed_obs (ref_frame indexes ref_pyramids). Returns (results array, summary)."""
    lib = load()
    S = len(seeds)
    T, Tk, frames, curp, lw, lh, sxp, syp, keep = _reproj_args(T_cur_w, T_f_w, ref_pyramids, cur_levels, cur_sobel)
    order = np.ascontiguousarray(cell_order, np.int32)
    out = (orc_reproj_result * max(S, 1))()
    summ = orc_reproj_summary()
    lib.orc_reproject_seeds(C.byref(cam_of(cam)), dp(T), Tk.size // 12, dp(Tk), S, seeds, C.byref(grid), order.ctypes.data_as(C.POINTER(C.c_int32)),
                            int(n_matches_in), int(max_search_level), frames, curp, lw, lh, sxp, syp, out, C.byref(summ))
    return out, summ


def reproject_seeds_speculative(cam, T_cur_w, T_f_w, seeds, grid, max_search_level, ref_pyramids, cur_levels, cur_sobel):
    lib = load()
    S = len(seeds)
    T, Tk, frames, curp, lw, lh, sxp, syp, keep = _reproj_args(T_cur_w, T_f_w, ref_pyramids, cur_levels, cur_sobel)
    out = (orc_reproj_result * max(S, 1))()
    px_after = np.zeros((max(S, 1), 2))
    lib.orc_reproject_seeds_speculative(C.byref(cam_of(cam)), dp(T), Tk.size // 12, dp(Tk), S, seeds, C.byref(grid), int(max_search_level), frames, curp,
                                        lw, lh, sxp, syp, out, dp(px_after))
    return out, px_after


def seed_select(seeds, match_ok, grid, cell_order, n_matches_in, io):
    lib = load()
    ok = np.ascontiguousarray(match_ok, np.uint8)
    order = np.ascontiguousarray(cell_order, np.int32)
    summ = orc_reproj_summary()
    lib.orc_seed_select(len(seeds), seeds, ok.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(grid), order.ctypes.data_as(C.POINTER(C.c_int32)),
                        int(n_matches_in), io, C.byref(summ))
    return summ
