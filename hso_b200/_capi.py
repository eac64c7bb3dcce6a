"""ctypes binding of the C-ABI in include/hso_b200.h (libhso_b200.so, built in-tree by hso_b200/csrc/Makefile).

There is no CPU fallback: if the shared library is missing this module raises at import of the symbol table, and
`hso_create` fails with HSO_ERR_NO_DEVICE on a machine without an sm_100 GPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhso_b200.so")

HSO_OK = 0
HSO_ERR_INVALID = -1
HSO_ERR_CUDA = -2
HSO_ERR_NO_DEVICE = -3
HSO_ERR_CAPACITY = -4
HSO_ERR_BAD_FRAME = -5


class hso_cam(C.Structure):
    _fields_ = [("model", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("undistort", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("d", C.c_double * 5)]


class hso_cfg(C.Structure):
    _fields_ = [("n_pyr_levels", C.c_int32), ("klt_max_level", C.c_int32), ("max_frames", C.c_int32), ("max_features", C.c_int32),
                ("materialize_sobel", C.c_int32), ("reserved", C.c_int32 * 3)]


class hso_track_params(C.Structure):
    _fields_ = [("inverse_comp", C.c_int32), ("max_level", C.c_int32), ("min_level", C.c_int32), ("n_iter", C.c_int32)]


class hso_track_job(C.Structure):
    _fields_ = [("ref", C.c_int32), ("cur", C.c_int32), ("n_features", C.c_int32), ("reserved", C.c_int32),
                ("px", C.POINTER(C.c_double)), ("f", C.POINTER(C.c_double)), ("dist", C.POINTER(C.c_double)),
                ("T_cur_ref", C.c_double * 12), ("exposure_rat", C.c_float), ("reserved2", C.c_float),
                ("xyz", C.POINTER(C.c_double)), ("px32", C.POINTER(C.c_float))]


class hso_track_result(C.Structure):
    _fields_ = [("T_cur_ref", C.c_double * 12), ("exposure_rat", C.c_float), ("n_iters", C.c_int32), ("n_evals", C.c_int32),
                ("iters_per_level", C.c_int32 * 8), ("n_tracked", C.c_uint64), ("visible_patch_evals", C.c_uint64 * 8),
                ("trace_len", C.c_int32), ("reserved", C.c_int32), ("cycles", C.c_uint64 * 8)]


class hso_trace(C.Structure):
    _fields_ = [("level", C.c_int32), ("iter", C.c_int32), ("T_eval", C.c_double * 12), ("a_eval", C.c_float), ("lambda_", C.c_float),
                ("H", C.c_double * 49), ("b", C.c_double * 7), ("step", C.c_double * 7), ("energy", C.c_double),
                ("total_terms", C.c_int32), ("saturated_terms", C.c_int32), ("accepted", C.c_int32),
                ("huber", C.c_float), ("outlier", C.c_float)]


class hso_align_job(C.Structure):
    _fields_ = [("ref_level", C.c_int32), ("search_level", C.c_int32), ("type", C.c_int32), ("scale_patch", C.c_int32),
                ("px_ref", C.c_double * 2), ("A_cur_ref", C.c_double * 4), ("grad", C.c_double * 2), ("px_cur", C.c_double * 2),
                ("exposure_rat", C.c_float), ("ncc_thresh", C.c_float)]


class hso_align_result(C.Structure):
    _fields_ = [("ok", C.c_int32), ("align_converged", C.c_int32), ("px_cur", C.c_double * 2), ("h_inv", C.c_double)]


class hso_reproj_cand(C.Structure):
    _fields_ = [("p_host", C.c_double * 3), ("px_ref", C.c_double * 2), ("f_ref", C.c_double * 3), ("grad", C.c_double * 2),
                ("depth_ref", C.c_double), ("host_pose", C.c_int32), ("ref_pose", C.c_int32), ("ref_frame", C.c_int32),
                ("ref_level", C.c_int32), ("ftr_type", C.c_int32), ("pt_type", C.c_int32), ("pt_ftr_type", C.c_int32),
                ("scale_patch", C.c_int32), ("exposure_rat", C.c_float), ("pad_", C.c_float)]


class hso_reproj_grid(C.Structure):
    _fields_ = [("cell_size", C.c_int32), ("n_cols", C.c_int32), ("n_rows", C.c_int32), ("max_fts", C.c_int32),
                ("align_max_iter", C.c_int32), ("pad_", C.c_int32)]


class hso_reproj_result(C.Structure):
    _fields_ = [("in_frame", C.c_int32), ("cell", C.c_int32), ("tried", C.c_int32), ("matched", C.c_int32), ("search_level", C.c_int32),
                ("order", C.c_int32), ("align_ok", C.c_int32), ("pad_", C.c_int32), ("px", C.c_double * 2), ("A_cur_ref", C.c_double * 4)]


class hso_reproj_summary(C.Structure):
    _fields_ = [("n_in_frame", C.c_int32), ("n_matches", C.c_int32), ("n_trials", C.c_int32), ("used_cell_all", C.c_int32)]


class hso_seed_obs(C.Structure):
    _fields_ = [("px", C.c_double * 2), ("f", C.c_double * 3), ("grad", C.c_double * 2), ("ref_frame", C.c_int32), ("ref_pose", C.c_int32),
                ("level", C.c_int32), ("ftr_type", C.c_int32), ("mu", C.c_float), ("sigma2", C.c_float), ("exposure_rat", C.c_float),
                ("pad_", C.c_float)]


class hso_seed_result(C.Structure):
    _fields_ = [("is_update", C.c_int32), ("is_valid", C.c_int32), ("res", C.c_int32), ("search_level", C.c_int32), ("epl_start", C.c_int32 * 2),
                ("epl_end", C.c_int32 * 2), ("mu", C.c_float), ("sigma2", C.c_float), ("z", C.c_double), ("px_cur", C.c_double * 2)]


class hso_corner(C.Structure):
    _fields_ = [("x", C.c_int16), ("y", C.c_int16), ("score", C.c_int32), ("shi_tomasi", C.c_float)]


class hso_pose_result(C.Structure):
    _fields_ = [("T_f_w", C.c_double * 12), ("cov", C.c_double * 36), ("estimated_scale", C.c_double), ("error_init", C.c_double),
                ("error_final", C.c_double), ("num_obs", C.c_uint64), ("error_in_px", C.c_float), ("n_trials_total", C.c_int32),
                ("early_return", C.c_int32)]


# every symbol include/hso_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
_vp = C.c_void_p
SYMBOLS = {
    "hso_cfg_default": (None, [_P(hso_cfg)]),
    "hso_create": (C.c_int, [C.c_int, _P(hso_cam), _P(hso_cfg), _P(_vp)]),
    "hso_destroy": (None, [_vp]),
    "hso_last_error": (C.c_char_p, [_vp]),
    "hso_set_stream": (C.c_int, [_vp, _vp]),
    "hso_get_stream": (_vp, [_vp]),
    "hso_synchronize": (C.c_int, [_vp]),
    "hso_kernel_launches": (C.c_uint64, [_vp]),
    "hso_frame_upload": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _P(C.c_int32), _P(C.c_float), _P(C.c_float)]),
    "hso_frame_upload_batch": (C.c_int, [_vp, C.c_int, _P(_vp), C.c_int, C.c_int, C.c_int, _P(C.c_int32), _P(C.c_float), _P(C.c_float)]),
    "hso_frame_upload_raw_batch": (C.c_int, [_vp, C.c_int, _P(_vp), C.c_int, C.c_int, C.c_int, C.c_int, _P(C.c_int32), _P(C.c_float), _P(C.c_float)]),
    "hso_undistort_maps": (C.c_int, [_vp, _vp, _vp]),
    "hso_frame_build_batch_device": (C.c_int, [_vp, C.c_int, _P(_vp), C.c_int, C.c_int, C.c_int, _P(C.c_int32)]),
    "hso_frame_rebuild_batch_device": (C.c_int, [_vp, C.c_int, _P(_vp), C.c_int, C.c_int, C.c_int, _P(C.c_int32)]),
    "hso_frame_stats": (C.c_int, [_vp, C.c_int32, _P(C.c_float), _P(C.c_float)]),
    "hso_frame_level_size": (C.c_int, [_vp, C.c_int32, C.c_int, _P(C.c_int), _P(C.c_int)]),
    "hso_frame_download_level": (C.c_int, [_vp, C.c_int32, C.c_int, _vp]),
    "hso_frame_download_sobel": (C.c_int, [_vp, C.c_int32, C.c_int, _vp, _vp]),
    "hso_frame_release": (C.c_int, [_vp, C.c_int32]),
    "hso_frame_release_batch": (C.c_int, [_vp, C.c_int, _P(C.c_int32)]),
    "hso_coarse_track": (C.c_int, [_vp, _P(hso_track_params), _P(hso_track_job), _P(hso_track_result), _P(hso_trace), C.c_int, _P(C.c_int)]),
    "hso_coarse_track_batch": (C.c_int, [_vp, _P(hso_track_params), C.c_int, _P(hso_track_job), _P(hso_track_result), _P(hso_trace), C.c_int,
                                         _P(C.c_int)]),
    "hso_add_frames_track_batch": (C.c_int, [_vp, _P(hso_track_params), C.c_int, _P(_vp), C.c_int, C.c_int, C.c_int, _P(hso_track_job), _P(C.c_int32),
                                             _P(C.c_float), _P(C.c_float), _P(hso_track_result)]),
    "hso_set_pipeline": (C.c_int, [_vp, C.c_int, C.c_int]),
    "hso_track_stage": (C.c_int, [_vp, _P(hso_track_params), C.c_int, _P(hso_track_job), C.c_int]),
    "hso_track_restage_frames": (C.c_int, [_vp, C.c_int, _P(C.c_int32), _P(C.c_int32)]),
    "hso_track_run": (C.c_int, [_vp]),
    "hso_track_collect": (C.c_int, [_vp, _P(hso_track_result), _P(hso_trace), _P(C.c_int)]),
    "hso_track_set_profile": (C.c_int, [_vp, C.c_int]),
    "hso_track_level_profile": (C.c_int, [_vp, C.c_int, _P(C.c_double), _P(C.c_uint64)]),
    "hso_track_set_cluster": (C.c_int, [_vp, C.c_int, C.c_int]),
    "hso_track_set_level_shape": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "hso_track_get_level_shape": (C.c_int, [_vp, C.c_int, _P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "hso_track_set_ic_dual": (C.c_int, [_vp, C.c_int]),
    "hso_track_set_stream_cache": (C.c_int, [_vp, C.c_int]),
    "hso_track_set_direct_inputs": (C.c_int, [_vp, C.c_int]),
    "hso_align_batch": (C.c_int, [_vp, C.c_int32, C.c_int, _P(hso_align_job), _P(C.c_int32), C.c_int, _P(hso_align_result)]),
    "hso_reproject_match": (C.c_int, [_vp, C.c_int32, _P(C.c_double), C.c_int, _P(C.c_double), C.c_int, _P(hso_reproj_cand), _P(hso_reproj_grid),
                                      _P(C.c_int32), _P(hso_reproj_result), _P(hso_reproj_summary)]),
    "hso_reproject_select_only": (C.c_int, [_vp, C.c_int, _P(hso_reproj_cand), _P(C.c_int32), _P(C.c_int32), _P(C.c_uint8), _P(hso_reproj_grid),
                                            _P(C.c_int32), _P(hso_reproj_result), _P(hso_reproj_summary)]),
    "hso_reproject_seeds": (C.c_int, [_vp, C.c_int32, _P(C.c_double), C.c_int, _P(C.c_double), C.c_int, _P(hso_seed_obs), _P(hso_reproj_grid), _P(C.c_int32),
                                      C.c_int, _P(hso_reproj_result), _P(hso_reproj_summary)]),
    "hso_depth_observe": (C.c_int, [_vp, C.c_int32, _P(C.c_double), C.c_int, _P(C.c_double), C.c_double, C.c_int, C.c_int, _P(hso_seed_obs),
                                    _P(hso_seed_result)]),
    "hso_pose_optimize": (C.c_int, [_vp, C.c_double, C.c_int, C.c_int, C.c_int, _P(C.c_double), _P(C.c_double), _P(C.c_int32), C.c_int,
                                    _P(C.c_double), _P(C.c_double), _P(C.c_int8), _P(C.c_int8), _P(C.c_int8), _P(C.c_double),
                                    _P(C.c_uint8), _P(hso_pose_result)]),
    "hso_pose_optimize_batch": (C.c_int, [_vp, C.c_double, C.c_int, C.c_int, _P(C.c_int32), _P(C.c_int32), _P(C.c_double), _P(C.c_double),
                                          _P(C.c_int32), _P(C.c_int32), _P(C.c_double), _P(C.c_double), _P(C.c_int8), _P(C.c_int8),
                                          _P(C.c_int8), _P(C.c_double), _P(C.c_uint8), _P(hso_pose_result)]),
    "hso_fast_detect": (C.c_int, [_vp, C.c_int32, C.c_int, C.c_int, C.c_int, _P(hso_corner), C.c_int, _P(C.c_int)]),
    "hso_fast_detect_levels": (C.c_int, [_vp, C.c_int32, C.c_int, C.c_int, C.c_int, _P(hso_corner), C.c_int, _P(C.c_int)]),
    "hso_stage_name": (C.c_char_p, [C.c_int]),
    "hso_stage_time_ms": (C.c_int, [_vp, C.c_int, _P(C.c_double), _P(C.c_uint64)]),
}

_lib = None


def load():
    """Load libhso_b200.so and bind every declared symbol. Raises (loudly) if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(make -C hso_b200/csrc). hso_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
