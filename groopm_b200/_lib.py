"""ctypes binding of libgmdist.so (the C ABI declared in include/gmdist.h).

There is no CPU fallback: if the shared library has not been built, importing the
engine fails loudly; if no CUDA device is present every compute call fails with
the CUDA error text from the library.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# GMDIST_LIB: developer override to A/B another build of the same ABI on one GPU box
LIB_PATH = Path(os.environ.get("GMDIST_LIB") or Path(__file__).resolve().parent / "libgmdist.so")

OK, ERR_BAD_ARG, ERR_NONFINITE, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
L2, L1 = 0, 1
FORM_EXPANDED, FORM_DIRECT_L2, FORM_DIRECT_L1, FORM_TC, FORM_AUTO = 0, 1, 2, 3, -1
TILE_ROWS, MAX_K, MAX_D = 64, 64, 136

_i64, _int, _vp, _sz = C.c_int64, C.c_int, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); every symbol include/gmdist.h declares
SIGNATURES = {
    "gmd_last_error": (C.c_char_p, []),
    "gmd_version": (_int, []),
    "gmd_bmu_host_f64": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp]),
    "gmd_knn_host_f64": (_int, [_vp, _i64, _vp, _i64, _int, _int, _int, _vp, _vp]),
    "gmd_tiles_bytes": (_sz, [_i64, _int]),
    "gmd_col_mean_f64": (_int, [_vp, _i64, _int, _vp, _vp]),
    "gmd_prep_f64": (_int, [_vp, _i64, _int, _vp, _int, _vp, _vp, _vp, _vp]),
    "gmd_scan_topk_f32": (_int, [_vp, _i64, _vp, _i64, _int, _int, _int, _int, _i64, _vp, _vp, _vp, _vp]),
    "gmd_use_bmu_scan": (_int, [_int, _int, _int]),
    "gmd_bmu_kk": (_int, []),
    "gmd_scan_bmu_f32": (_int, [_vp, _i64, _vp, _i64, _int, _i64, _vp, _vp, _vp, _vp]),
    "gmd_tc_supported": (_int, [_int, _int, _int]),
    "gmd_use_tc_scan": (_int, [_i64, _i64, _int, _int, _int]),
    "gmd_tc_tiles_bytes": (_sz, [_i64, _int]),
    "gmd_tc_norms_f64": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _vp]),
    "gmd_prep_tc_f64": (_int, [_vp, _i64, _int, _vp, _int, _vp, _vp, _vp, _vp]),
    "gmd_scan_tc_f32": (_int, [_vp, _i64, _vp, _i64, _int, _int, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gmd_refine_topk_f64": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _int, _int, _int, _int, _int,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gmd_refine_topk_order_f64": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _int, _int, _int, _int, _int,
                                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "gmd_prep_tc_perm_f64": (_int, [_vp, _i64, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "gmd_scan_tc_sweep_f32": (_int, [_vp, _i64, _vp, _i64, _int, _int, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp]),
    "gmd_group_shells": (_int, []),
    "gmd_group_params": (None, [_i64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gmd_group_positions": (_i64, [_i64, _int, _int]),
    "gmd_group_leaders_workspace_bytes": (_sz, [_int, _int]),
    "gmd_group_leaders_f64": (_int, [_vp, _i64, _int, _vp, _int, _int, C.c_float, _vp, _vp, _vp, _vp, _vp, _sz,
                                     C.POINTER(C.c_int32), _vp]),
    "gmd_group_order_workspace_bytes": (_sz, [_int]),
    "gmd_group_order": (_int, [_vp, _vp, _i64, _int, _vp, _int, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "gmd_group_order_pad": (_int, [_vp, _vp, _i64, _int, _vp, _int, _int, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "gmd_prep_tc_grouped_f64": (_int, [_vp, _i64, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "gmd_prep_tc_grouped_both_f64": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "gmd_tc_group_meta": (_int, [_vp, _vp, _int, _int, _vp, _vp, _vp, _int, _vp]),
    "gmd_use_tc_prune": (_int, [_i64, _i64, _int, _int, _int]),
    "gmd_tc_prune_max_tiles": (_sz, [_int, _int]),
    "gmd_tc_grouped_tiles_bytes": (_sz, [_i64, _i64, _int]),
    "gmd_tc_grouped_workspace_bytes": (_sz, [_i64, _i64, _int]),
    "gmd_tc_stage_grouped_f64": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gmd_tc_grouped_leaders_f64": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _sz, C.POINTER(C.c_int32),
                                          C.POINTER(_vp), _vp]),
    "gmd_tc_stage_grouped_ext_f64": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, C.c_int32,
                                            _vp, _vp, _vp, _vp]),
    "gmd_assign_leaders_f64": (_int, [_vp, _i64, _vp, _int, _int, _vp, _vp, _vp, _sz, _vp]),
    "gmd_pick_proposal_f64": (_int, [_vp, _i64, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "gmd_tc_restage_queries_f64": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gmd_scan_tc_prune_f32": (_int, [_vp, _i64, _vp, _i64, _int, _int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp,
                                     _int, _vp, _int, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "gmd_scan_tc_grouped_f32": (_int, [_vp, _vp, _int, _int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gmd_order_workspace_bytes": (_sz, [_int]),
    "gmd_order_by_key": (_int, [_vp, _i64, _int, _vp, _vp, _sz, _vp]),
    "gmd_exact_workspace_bytes": (_sz, [_i64]),
    "gmd_exact_rows_f64": (_int, [_vp, _i64, _vp, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gmd_knn_workspace_bytes": (_sz, [_i64, _i64, _int, _int]),
    "gmd_knn_f64": (_int, [_vp, _i64, _vp, _i64, _int, _int, _int, _int, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gmd_cdist_f64": (_int, [_vp, _i64, _vp, _i64, _int, _int, _vp, _vp]),
    "gmd_angle_block_f64": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp]),
    "gmd_pdist_f64": (_int, [_vp, _i64, _int, _int, _vp, _vp]),
    "gmd_radius_count_f64": (_int, [_vp, _i64, _int, _int, C.c_double, _vp, _vp]),
    "gmd_radius_fill_f64": (_int, [_vp, _i64, _int, _int, C.c_double, _vp, _vp, _vp, _vp]),
    "gmd_gather_dist_f64": (_int, [_vp, _i64, _int, _int, _vp, _vp, _vp, _vp]),
    "gmd_median_workspace_bytes": (_sz, []),
    "gmd_median_dist_f64": (_int, [_vp, _i64, _vp, _i64, _int, _int, _vp, _vp, _vp]),
    "gmd_median_batch_f64": (_int, [_vp, _i64, _int, _int, _vp, _vp, _int, _vp, _vp]),
    "gmd_nonfinite_f64": (_int, [_vp, _i64, _vp, _vp]),
    "gmd_zero_async": (_int, [_vp, _sz, _vp]),
    "gmd_bmu_small_f64": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp]),
    "gmd_knn_small_max_candidates": (_int, []),
    "gmd_knn_small_f64": (_int, [_vp, _i64, _vp, _i64, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "gmd_indptr_from_counts_i64": (_int, [_vp, _i64, _vp, _vp]),
    "gmd_som_update_f64": (_int, [_vp, _vp, _int, _int, _int, _vp, _i64, _vp, _vp, _int, _vp, _int, _vp, _vp]),
    "gmd_pick_form": (_int, [_int, _int]),
    "gmd_pick_kk": (_int, [_int]),
    "gmd_pick_nsplit": (_int, [_i64, _i64, _int, _int]),
    "gmd_fp32_peak_probe": (_int, [_int, _vp, C.POINTER(C.c_double), _vp]),
}


class Grouped(C.Structure):
    """gmd_grouped (include/gmdist.h): what the pruned tensor-core scan needs from its staging."""
    _fields_ = [("npos_q", C.c_int64), ("npos_c", C.c_int64), ("ngroups", C.c_int32), ("ng_alloc", C.c_int32),
                ("nt_alloc_q", C.c_int32), ("nt_alloc_c", C.c_int32), ("same", C.c_int32), ("reserved", C.c_int32),
                ("perm_q", _vp), ("perm_c", _vp), ("tgroup_q", _vp), ("gstart_c", _vp), ("gmeta", _vp),
                ("tmeta_q", _vp), ("tmeta_c", _vp), ("counters", _vp)]


class GmdistError(RuntimeError):
    pass


def load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m groopm_b200.build` "
            "(there is no CPU fallback for the distance engine)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


def check(status: int) -> None:
    """Map a gmd_status to the exception the reference's call site would see."""
    if status == OK:
        return
    msg = lib().gmd_last_error().decode("utf-8", "replace")
    if status == ERR_NONFINITE:
        # the reference runs under np.seterr(all='raise') (torusMesh.py:57, som.py:80)
        raise FloatingPointError(msg)
    if status in (ERR_BAD_ARG, ERR_WORKSPACE, ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise GmdistError(msg)
