"""Host-side (numpy) mirror of the reference interfaces of the tracking hot path, on top of the C-ABI.

Names follow the reference: Frame (include/hso/frame.h:131), CoarseTracker(inverse_composition, max_level, min_level, n_iter).run
(include/hso/CoarseTracker.h:134,141), pose_optimizer.optimizeLevenbergMarquardt3rd (include/hso/pose_optimizer.h:61-64),
Matcher.findMatchDirect's inner part as align_batch (include/hso/matcher.h:153). All compute runs in libhso_b200.so on the GPU;
this module only flattens arguments. No oracle/, no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _capi as K

PINHOLE, FOV, EQUIDISTANT = 0, 1, 2


class HsoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hso_b200 error {code}: {msg}")
        self.code = code


def make_cam(width, height, fx, fy, cx, cy, d=(0, 0, 0, 0, 0), model=PINHOLE, undistort=0):
    c = K.hso_cam()
    c.model, c.width, c.height, c.undistort = model, width, height, undistort
    c.fx, c.fy, c.cx, c.cy = fx, fy, cx, cy
    for i in range(5):
        c.d[i] = float(d[i]) if i < len(d) else 0.0
    return c


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """One hso_ctx: one CUDA stream, one camera model, a table of device-resident frames."""

    def __init__(self, cam, device=0, max_frames=16, max_features=8192, materialize_sobel=False, n_pyr_levels=3, klt_max_level=4):
        self.lib = K.load()
        cfg = K.hso_cfg()
        self.lib.hso_cfg_default(C.byref(cfg))
        cfg.max_frames, cfg.max_features = max_frames, max_features
        cfg.materialize_sobel = 1 if materialize_sobel else 0
        cfg.n_pyr_levels, cfg.klt_max_level = n_pyr_levels, klt_max_level
        self.cam = cam
        self.h = C.c_void_p()
        rc = self.lib.hso_create(device, C.byref(cam), C.byref(cfg), C.byref(self.h))
        if rc != K.HSO_OK:
            self.h = None
            raise HsoError(rc, "hso_create failed (no sm_100 CUDA device? hso_b200 has no CPU fallback)")
        self.n_levels = max(n_pyr_levels, klt_max_level + 1)
        self._keep = []

    def close(self):
        if self.h:
            self.lib.hso_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != K.HSO_OK:
            raise HsoError(rc, self.lib.hso_last_error(self.h).decode())

    # ---- F1: Frame ------------------------------------------------------------------------------------------------------------
    def upload_frames(self, imgs):
        """imgs: list of HxW uint8 arrays (C-contiguous rows; arbitrary row stride allowed). Returns (ids, integral, grad_mean)."""
        B = len(imgs)
        H, W = imgs[0].shape
        stride = imgs[0].strides[0]
        ptrs = (C.c_void_p * B)(*[im.ctypes.data for im in imgs])
        ids = (C.c_int32 * B)()
        integral = np.zeros(B, np.float32)
        gm = np.zeros(B, np.float32)
        self._chk(self.lib.hso_frame_upload_batch(self.h, B, ptrs, W, H, stride, ids,
                                                  integral.ctypes.data_as(C.POINTER(C.c_float)), gm.ctypes.data_as(C.POINTER(C.c_float))))
        return list(ids), integral, gm

    def upload_raw_frames(self, imgs, undistort=False):
        """Raw images (any size; resized to the camera size like ImageReader::readImage when it differs) -> optional undistortion remap
        (AbstractCamera::undistortImage) -> Frame. Returns (ids, integral, grad_mean)."""
        B = len(imgs)
        imgs = [im if (im.dtype == np.uint8 and im.strides[1] == 1) else np.ascontiguousarray(im, np.uint8) for im in imgs]  # row padding is allowed
        H, W = imgs[0].shape
        ptrs = (C.c_void_p * B)(*[im.ctypes.data for im in imgs])
        ids = (C.c_int32 * B)()
        integral = np.zeros(B, np.float32)
        gm = np.zeros(B, np.float32)
        fp = C.POINTER(C.c_float)
        self._chk(self.lib.hso_frame_upload_raw_batch(self.h, B, ptrs, W, H, imgs[0].strides[0], int(bool(undistort)), ids,
                                                      integral.ctypes.data_as(fp), gm.ctypes.data_as(fp)))
        return list(ids), integral, gm

    def undistort_maps(self):
        w, h = self.level_size(0)
        m1 = np.zeros((h, w, 2), np.int16)
        m2 = np.zeros((h, w), np.uint16)
        self._chk(self.lib.hso_undistort_maps(self.h, m1.ctypes.data, m2.ctypes.data))
        return m1, m2

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        self._chk(self.lib.hso_frame_level_size(self.h, 0, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def download_level(self, fid, level):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        self._chk(self.lib.hso_frame_download_level(self.h, fid, level, out.ctypes.data))
        return out

    def download_sobel(self, fid, level):
        w, h = self.level_size(level)
        gx = np.empty((h, w), np.int16)
        gy = np.empty((h, w), np.int16)
        self._chk(self.lib.hso_frame_download_sobel(self.h, fid, level, gx.ctypes.data, gy.ctypes.data))
        return gx, gy

    def release(self, fid):
        self._chk(self.lib.hso_frame_release(self.h, fid))

    # ---- F2: CoarseTracker ----------------------------------------------------------------------------------------------------
    @staticmethod
    def compact_features(px, f, dist):
        """The compact input layout of hso_track_job, as a caller's gather loop over ref_frame->fts_ forms it: features without a point / with a
        negative distance left out, xyz = f * dist (Vector3d xyz_ref((*it_ft)->f*dist), src/CoarseTracker.cpp:292), px as float32."""
        dist = np.asarray(dist, np.float64)
        ok = dist >= 0
        xyz = np.ascontiguousarray(np.asarray(f, np.float64)[ok] * dist[ok, None])
        return xyz, np.ascontiguousarray(np.asarray(px, np.float64)[ok].astype(np.float32))

    def _track_jobs(self, jobs):
        B = len(jobs)
        arr = (K.hso_track_job * B)()
        keep = []
        for b, j in enumerate(jobs):
            a = arr[b]
            if "xyz" in j:  # compact layout (hso_track_job::xyz / px32, see compact_features)
                xyz = np.ascontiguousarray(j["xyz"], np.float64).reshape(-1)
                px32 = np.ascontiguousarray(j["px32"], np.float32).reshape(-1)
                keep += [xyz, px32]
                a.ref, a.cur, a.n_features = int(j["ref"]), int(j["cur"]), px32.shape[0] // 2
                a.xyz, a.px32 = _dp(xyz), px32.ctypes.data_as(C.POINTER(C.c_float))
            else:
                px = np.ascontiguousarray(j["px"], np.float64).reshape(-1)
                f = np.ascontiguousarray(j["f"], np.float64).reshape(-1)
                dist = np.ascontiguousarray(j["dist"], np.float64).reshape(-1)
                keep += [px, f, dist]
                a.ref, a.cur, a.n_features = int(j["ref"]), int(j["cur"]), dist.shape[0]
                a.px, a.f, a.dist = _dp(px), _dp(f), _dp(dist)
            T = np.ascontiguousarray(j["T_cur_ref"], np.float64).reshape(12)
            for k in range(12):
                a.T_cur_ref[k] = T[k]
            a.exposure_rat = float(j["exposure_rat"])
        return arr, keep

    def coarse_track_batch(self, jobs, inverse_comp=False, max_level=4, min_level=1, n_iter=50, trace_cap=0):
        """jobs: list of dicts {ref, cur, px (F,2), f (F,3), dist (F,), T_cur_ref (3,4), exposure_rat}. Returns (results, traces)."""
        B = len(jobs)
        prm = K.hso_track_params(int(bool(inverse_comp)), max_level, min_level, n_iter)
        arr, keep = self._track_jobs(jobs)
        out = (K.hso_track_result * B)()
        trace = (K.hso_trace * (B * trace_cap))() if trace_cap else None
        tlen = (C.c_int * B)()
        self._chk(self.lib.hso_coarse_track_batch(self.h, C.byref(prm), B, arr, out, trace, trace_cap, tlen))
        res = []
        for b in range(B):
            o = out[b]
            res.append(dict(T_cur_ref=np.array(o.T_cur_ref[:]).reshape(3, 4), exposure_rat=float(o.exposure_rat), n_iters=o.n_iters,
                            n_evals=o.n_evals, iters_per_level=list(o.iters_per_level), n_tracked=int(o.n_tracked),
                            visible_patch_evals=list(o.visible_patch_evals)))
        traces = []
        for b in range(B):
            traces.append([trace[b * trace_cap + i] for i in range(tlen[b])] if trace_cap else [])
        return res, traces

    def add_frames_track_batch(self, images, jobs, inverse_comp=False, max_level=4, min_level=1, n_iter=50):
        """Frame construction + CoarseTracker::run for B independent (image, reference frame) pairs in one chunk-pipelined call
        (FrameHandlerMono::addImage's front end). jobs as in coarse_track_batch without 'cur'; 'exposure_rat' < 0 (default) lets the
        device form cur.integralImage_/ref.integralImage_. Returns (new frame ids, integral, grad_mean, results)."""
        B = len(jobs)
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        H, W = imgs[0].shape
        ptrs = (C.c_void_p * B)(*[im.ctypes.data for im in imgs])
        prm = K.hso_track_params(int(bool(inverse_comp)), max_level, min_level, n_iter)
        arr, keep = self._track_jobs([dict(j, cur=j.get("cur", 0), exposure_rat=j.get("exposure_rat", -1.0)) for j in jobs])
        ids = (C.c_int32 * B)()
        integ = np.zeros(B, np.float32)
        gm = np.zeros(B, np.float32)
        out = (K.hso_track_result * B)()
        fp = C.POINTER(C.c_float)
        self._chk(self.lib.hso_add_frames_track_batch(self.h, C.byref(prm), B, ptrs, W, H, W, arr, ids, integ.ctypes.data_as(fp),
                                                      gm.ctypes.data_as(fp), out))
        res = [dict(T_cur_ref=np.array(o.T_cur_ref[:]).reshape(3, 4), exposure_rat=float(o.exposure_rat), n_iters=o.n_iters, n_evals=o.n_evals,
                    iters_per_level=list(o.iters_per_level), n_tracked=int(o.n_tracked)) for o in out]
        return list(ids), integ, gm, res

    def track_stage(self, jobs, inverse_comp=False, max_level=4, min_level=1, n_iter=50):
        prm = K.hso_track_params(int(bool(inverse_comp)), max_level, min_level, n_iter)
        arr, keep = self._track_jobs(jobs)
        self._chk(self.lib.hso_track_stage(self.h, C.byref(prm), len(jobs), arr, 0))
        self._tB = len(jobs)

    def track_run(self):
        self._chk(self.lib.hso_track_run(self.h))

    def track_collect(self):
        out = (K.hso_track_result * self._tB)()
        self._chk(self.lib.hso_track_collect(self.h, out, None, None))
        return out

    # ---- F3-inner: Matcher::findMatchDirect after getWarpMatrixAffine ------------------------------------------------------
    def align_batch(self, cur, jobs, ref_frames, align_max_iter=10):
        """jobs: ctypes array of hso_align_job (or list of dicts with its field names); ref_frames: frame id per job.
        Returns a ctypes array of hso_align_result."""
        M = len(jobs)
        if M == 0:
            self._chk(self.lib.hso_align_batch(self.h, int(cur), 0, None, None, align_max_iter, None))
            return []
        if not isinstance(jobs, C.Array):
            arr = (K.hso_align_job * M)()
            for m, j in enumerate(jobs):
                a = arr[m]
                a.ref_level, a.search_level, a.type, a.scale_patch = j["ref_level"], j["search_level"], j["type"], j.get("scale_patch", 0)
                for k in range(2):
                    a.px_ref[k], a.grad[k], a.px_cur[k] = j["px_ref"][k], j["grad"][k], j["px_cur"][k]
                for k in range(4):
                    a.A_cur_ref[k] = float(np.asarray(j["A_cur_ref"]).reshape(4)[k])
                a.exposure_rat = j.get("exposure_rat", 1.0)
            jobs = arr
        refs = (C.c_int32 * max(M, 1))(*[int(r) for r in ref_frames])
        out = (K.hso_align_result * max(M, 1))()
        self._chk(self.lib.hso_align_batch(self.h, int(cur), M, jobs, refs, align_max_iter, out))
        return out

    # ---- N1: Reprojector::reprojectMap data path ---------------------------------------------------------------------------
    @staticmethod
    def reproj_cands(cands, frame_ids=None):
        """list of dicts (hso_reproj_cand field names; 'ref_frame' indexes frame_ids when given) -> ctypes array."""
        arr = (K.hso_reproj_cand * max(len(cands), 1))()
        for i, c in enumerate(cands):
            a = arr[i]
            for k in range(3):
                a.p_host[k], a.f_ref[k] = float(c["p_host"][k]), float(c["f_ref"][k])
            for k in range(2):
                a.px_ref[k], a.grad[k] = float(c["px_ref"][k]), float(c["grad"][k])
            a.depth_ref = float(c["depth_ref"])
            a.host_pose, a.ref_pose = int(c["host_pose"]), int(c["ref_pose"])
            a.ref_frame = int(frame_ids[c["ref_frame"]]) if frame_ids is not None else int(c["ref_frame"])
            a.ref_level, a.ftr_type, a.pt_type, a.pt_ftr_type = int(c["ref_level"]), int(c["ftr_type"]), int(c["pt_type"]), int(c["pt_ftr_type"])
            a.scale_patch, a.exposure_rat = int(c.get("scale_patch", 0)), float(c.get("exposure_rat", 1.0))
        return arr

    def reproject_match(self, cur, T_cur_w, T_f_w, cands, grid, cell_order, M=None):
        """cands: ctypes array from reproj_cands(); grid: dict(cell_size, n_cols, n_rows, max_fts, align_max_iter).
        Returns (ctypes array of hso_reproj_result, hso_reproj_summary)."""
        M = len(cands) if M is None else M
        T = np.ascontiguousarray(T_cur_w, np.float64).reshape(12)
        Tk = np.ascontiguousarray(T_f_w, np.float64).reshape(-1)
        g = K.hso_reproj_grid(int(grid["cell_size"]), int(grid["n_cols"]), int(grid["n_rows"]), int(grid["max_fts"]),
                              int(grid.get("align_max_iter", 10)), 0)
        order = np.ascontiguousarray(cell_order, np.int32)
        out = (K.hso_reproj_result * max(M, 1))()
        summ = K.hso_reproj_summary()
        self._chk(self.lib.hso_reproject_match(self.h, int(cur), _dp(T), Tk.size // 12, _dp(Tk), M, cands, C.byref(g),
                                               order.ctypes.data_as(C.POINTER(C.c_int32)), out, C.byref(summ)))
        return out, summ

    def reproject_seeds(self, cur, T_cur_w, T_f_w, seeds, grid, cell_order, n_matches_in=0, S=None):
        """a13b: the seed stage of Reprojector::reprojectMap. seeds: ctypes array from seed_obs(). Returns (hso_reproj_result array, summary)."""
        S = len(seeds) if S is None else S
        T = np.ascontiguousarray(T_cur_w, np.float64).reshape(12)
        Tk = np.ascontiguousarray(T_f_w, np.float64).reshape(-1)
        g = K.hso_reproj_grid(int(grid["cell_size"]), int(grid["n_cols"]), int(grid["n_rows"]), int(grid["max_fts"]),
                              int(grid.get("align_max_iter", 10)), 0)
        order = np.ascontiguousarray(cell_order, np.int32)
        out = (K.hso_reproj_result * max(S, 1))()
        summ = K.hso_reproj_summary()
        self._chk(self.lib.hso_reproject_seeds(self.h, int(cur), _dp(T), Tk.size // 12, _dp(Tk), S, seeds, C.byref(g),
                                               order.ctypes.data_as(C.POINTER(C.c_int32)), int(n_matches_in), out, C.byref(summ)))
        return out, summ

    def reproject_select_only(self, cands, in_frame, cell, align_ok, grid, cell_order):
        """Test hook: the selection kernel alone on given per-candidate facts. Returns (results, summary)."""
        M = len(in_frame)
        g = K.hso_reproj_grid(int(grid["cell_size"]), int(grid["n_cols"]), int(grid["n_rows"]), int(grid["max_fts"]),
                              int(grid.get("align_max_iter", 10)), 0)
        inf = np.ascontiguousarray(in_frame, np.int32)
        cl = np.ascontiguousarray(cell, np.int32)
        ok = np.ascontiguousarray(align_ok, np.uint8)
        order = np.ascontiguousarray(cell_order, np.int32)
        out = (K.hso_reproj_result * max(M, 1))()
        summ = K.hso_reproj_summary()
        ip = C.POINTER(C.c_int32)
        self._chk(self.lib.hso_reproject_select_only(self.h, M, cands, inf.ctypes.data_as(ip), cl.ctypes.data_as(ip), ok.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                     C.byref(g), order.ctypes.data_as(ip), out, C.byref(summ)))
        return out, summ

    # ---- N3: DepthFilter::observeDepthRow -------------------------------------------------------------------------------------
    @staticmethod
    def seed_obs(seeds, frame_ids=None):
        """list of dicts (hso_seed_obs field names; 'ref_frame' indexes frame_ids when given) -> ctypes array."""
        arr = (K.hso_seed_obs * max(len(seeds), 1))()
        for i, s in enumerate(seeds):
            a = arr[i]
            for k in range(2):
                a.px[k], a.grad[k] = float(s["px"][k]), float(s["grad"][k])
            for k in range(3):
                a.f[k] = float(s["f"][k])
            a.ref_frame = int(frame_ids[s["ref_frame"]]) if frame_ids is not None else int(s["ref_frame"])
            a.ref_pose, a.level, a.ftr_type = int(s["ref_pose"]), int(s["level"]), int(s["ftr_type"])
            a.mu, a.sigma2, a.exposure_rat = float(s["mu"]), float(s["sigma2"]), float(s.get("exposure_rat", 1.0))
        return arr

    def depth_observe(self, cur, T_cur_w, T_f_w, seeds, px_error_angle, align_max_iter=10, S=None):
        """seeds: ctypes array from seed_obs(). Returns a ctypes array of hso_seed_result."""
        S = len(seeds) if S is None else S
        T = np.ascontiguousarray(T_cur_w, np.float64).reshape(12)
        Tk = np.ascontiguousarray(T_f_w, np.float64).reshape(-1)
        out = (K.hso_seed_result * max(S, 1))()
        self._chk(self.lib.hso_depth_observe(self.h, int(cur), _dp(T), Tk.size // 12, _dp(Tk), float(px_error_angle), int(align_max_iter), S, seeds, out))
        return out

    # ---- F4: pose_optimizer::optimizeLevenbergMarquardt3rd -----------------------------------------------------------------
    def pose_args(self, problems, reproj_thresh=2.0, n_iter=12):
        """Flattens the problems once and returns (call, (offs, outlier, results)): call() runs hso_pose_optimize_batch on those buffers
        (what a C++ caller holding flattened arrays does per frame)."""
        B = len(problems)
        cat = lambda key, dt, w: np.ascontiguousarray(np.concatenate([np.asarray(p[key], dt).reshape(-1, w) for p in problems]) if B else np.zeros((0, w), dt))
        f, ph, g = cat("f", np.float64, 3), cat("p_host", np.float64, 3), cat("grad", np.float64, 2)
        hi = cat("host_idx", np.int32, 1)
        lv, ft, pt = cat("level", np.int8, 1), cat("ftype", np.int8, 1), cat("ptype", np.int8, 1)
        Th = np.ascontiguousarray(np.concatenate([np.asarray(p["T_host_w"], np.float64).reshape(-1, 12) for p in problems]))
        T0 = np.ascontiguousarray(np.stack([np.asarray(p["T_f_w"], np.float64).reshape(12) for p in problems]))
        offs = np.zeros(B + 1, np.int32)
        hoffs = np.zeros(B + 1, np.int32)
        for b, p in enumerate(problems):
            offs[b + 1] = offs[b] + np.asarray(p["f"]).reshape(-1, 3).shape[0]
            hoffs[b + 1] = hoffs[b] + np.asarray(p["T_host_w"]).reshape(-1, 12).shape[0]
        nf = np.array([int(p.get("n_fts_total", np.asarray(p["f"]).reshape(-1, 3).shape[0])) for p in problems], np.int32)
        outl = np.zeros(max(int(offs[B]), 1), np.uint8)
        out = (K.hso_pose_result * B)()
        i32, i8, u8 = C.POINTER(C.c_int32), C.POINTER(C.c_int8), C.POINTER(C.c_uint8)
        keep = (f, ph, g, hi, lv, ft, pt, Th, T0, hoffs, nf)

        def call(_keep=keep):
            self._chk(self.lib.hso_pose_optimize_batch(self.h, reproj_thresh, n_iter, B, nf.ctypes.data_as(i32), offs.ctypes.data_as(i32), _dp(f), _dp(ph),
                                                       hi.ctypes.data_as(i32), hoffs.ctypes.data_as(i32), _dp(Th), _dp(g), lv.ctypes.data_as(i8),
                                                       ft.ctypes.data_as(i8), pt.ctypes.data_as(i8), _dp(T0), outl.ctypes.data_as(u8), out))
        return call, (offs, outl, out)

    def pose_optimize_batch(self, problems, reproj_thresh=2.0, n_iter=12):
        """problems: list of dicts {f (F,3), p_host (F,3), host_idx (F,), T_host_w (K,3,4), grad (F,2), level, ftype, ptype (F,),
        T_f_w (3,4), n_fts_total}. Returns list of dicts with the reference's outputs + outlier mask."""
        B = len(problems)
        call, (offs, outl, out) = self.pose_args(problems, reproj_thresh, n_iter)
        call()
        res = []
        for b in range(B):
            o = out[b]
            res.append(dict(T_f_w=np.array(o.T_f_w[:]).reshape(3, 4), cov=np.array(o.cov[:]).reshape(6, 6), estimated_scale=o.estimated_scale,
                            error_init=o.error_init, error_final=o.error_final, num_obs=int(o.num_obs), error_in_px=float(o.error_in_px),
                            n_trials_total=o.n_trials_total, early_return=o.early_return, outlier=outl[offs[b]:offs[b + 1]].copy()))
        return res

    # ---- N2: FeatureExtractor::fastDetectST detector part -----------------------------------------------------------------------
    def fast_detect(self, fid, level, threshold, border=8, cap=65536):
        """Returns [(x, y, score, shi_tomasi)] in raster order (level pixels), like fastDetectST before the cell bookkeeping."""
        out = (K.hso_corner * cap)()
        n = C.c_int()
        self._chk(self.lib.hso_fast_detect(self.h, int(fid), level, int(threshold), border, out, cap, C.byref(n)))
        if n.value > cap:
            return self.fast_detect(fid, level, threshold, border, n.value)
        return [(out[i].x, out[i].y, out[i].score, out[i].shi_tomasi) for i in range(n.value)]

    def fast_detect_levels(self, fid, n_levels, threshold, border=8, cap=32768):
        """Levels 0 .. n_levels-1 in one call. Returns a list (per level) of [(x, y, score, shi_tomasi)]."""
        out = (K.hso_corner * (cap * n_levels))()
        cnt = (C.c_int * n_levels)()
        self._chk(self.lib.hso_fast_detect_levels(self.h, int(fid), n_levels, int(threshold), border, out, cap, cnt))
        if max(cnt) > cap:
            return self.fast_detect_levels(fid, n_levels, threshold, border, max(cnt))
        return [[(out[l * cap + i].x, out[l * cap + i].y, out[l * cap + i].score, out[l * cap + i].shi_tomasi) for i in range(cnt[l])] for l in range(n_levels)]

    def stage_time_ms(self, stage):
        ms, calls = C.c_double(), C.c_uint64()
        self._chk(self.lib.hso_stage_time_ms(self.h, stage, C.byref(ms), C.byref(calls)))
        return ms.value, calls.value

    def set_cluster(self, ctas=0, threads=0):
        self._chk(self.lib.hso_track_set_cluster(self.h, ctas, threads))

    def set_level_shape(self, level, ctas=0, threads=0):
        self._chk(self.lib.hso_track_set_level_shape(self.h, level, ctas, threads))

    def level_shape(self, level):
        """(ctas per problem, threads per CTA, mode, absres_smem) of the last tracker run at `level`."""
        v = [C.c_int() for _ in range(4)]
        self._chk(self.lib.hso_track_get_level_shape(self.h, level, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def synchronize(self):
        self._chk(self.lib.hso_synchronize(self.h))

    def kernel_launches(self):
        return int(self.lib.hso_kernel_launches(self.h))

    def stream(self):
        return self.lib.hso_get_stream(self.h)


class CoarseTracker:
    """Mirror of hso::CoarseTracker (include/hso/CoarseTracker.h:134-143): ctor(inverse_composition, max_level, min_level, n_iter),
    run(ref, cur) -> number of tracked patches; updates cur['T_f_w'] and cur['exposure_time'] like src/CoarseTracker.cpp:198-202.

    Frames are dicts: {id, T_f_w (3x4), integral, exposure_time, px (F,2), f (F,3), dist (F,)} where dist is makeDepthRef's output
    (src/CoarseTracker.cpp:210-240; < 0 for features without point)."""

    def __init__(self, ctx, inverse_composition, max_level, min_level, n_iter, verbose=False):
        self.ctx, self.ic, self.max_level, self.min_level, self.n_iter = ctx, inverse_composition, max_level, min_level, n_iter

    def run(self, ref, cur):
        if len(ref["dist"]) == 0:
            return 0  # src/CoarseTracker.cpp:53
        T_ref, T_cur = _rt44(ref["T_f_w"]), _rt44(cur["T_f_w"])
        T_cur_ref = (T_cur @ np.linalg.inv(T_ref))[:3]
        job = dict(ref=ref["id"], cur=cur["id"], px=ref["px"], f=ref["f"], dist=ref["dist"], T_cur_ref=T_cur_ref,
                   exposure_rat=np.float32(cur["integral"]) / np.float32(ref["integral"]))
        res, _ = self.ctx.coarse_track_batch([job], self.ic, self.max_level, self.min_level, self.n_iter)
        r = res[0]
        cur["T_f_w"] = (_rt44(r["T_cur_ref"]) @ T_ref)[:3]
        a = r["exposure_rat"]
        cur["exposure_time"] = ref["exposure_time"] if 0.99 < a < 1.01 else a * ref["exposure_time"]
        return r["n_tracked"]


def _rt44(T):
    M = np.eye(4)
    M[:3] = np.asarray(T, np.float64).reshape(3, 4)
    return M
