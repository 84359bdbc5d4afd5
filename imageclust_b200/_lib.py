"""ctypes binding of ``libimageclust_b200.so`` (C ABI: ``include/imageclust_b200.h``).

This is the same boundary the Go/cgo shim binds.  Loading the
library needs no GPU (so CPU tests can check the exported symbols); every
computing entry point needs one -- there is no CPU fallback, ``ic_create`` fails
without a CUDA device and the Python layer raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# IMAGECLUST_B200_LIB: another build of the same library (A/B runs of kernel variants: scripts/build_variants.sh)
LIB_PATH = os.environ.get("IMAGECLUST_B200_LIB") or os.path.join(HERE, "libimageclust_b200.so")

IC_OK = 0
IC_ERR_TOO_FEW = -1
IC_ERR_UNSAT = -2
IC_ERR_BAD_ARG = -3
IC_ERR_CUDA = -4
IC_ERR_OOM = -5
IC_ERR_STATE = -6
IC_ERR_TIMEOUT = -7
IC_ERR_INTERNAL = -9

GRAM_TCGEN05_3XTF32 = 0
GRAM_EXACT_FP32 = 1
GRAM_TCGEN05_I8 = 2

# every symbol include/imageclust_b200.h declares
SYMBOLS = [
    "ic_create", "ic_destroy", "ic_last_error", "ic_pinned_alloc", "ic_pinned_free", "ic_set_option",
    "ic_optimal_clusters", "ic_cluster_with_constraints", "ic_load", "ic_load_device",
    "ic_initial_distances", "ic_set_matrix", "ic_nn_init", "ic_find_closest", "ic_merge_loop",
    "ic_run_resident", "ic_build_clusters", "ic_read_matrix", "ic_read_slots", "ic_get_merge_trace",
    "ic_get_stats", "ic_get_loop_profile", "ic_time_kernel",
    "ic_shard_init", "ic_shard_export", "ic_shard_connect", "ic_shard_rows",
    "ic_load_combined", "ic_read_x", "ic_get_loop_block_waits", "ic_get_linkage",
]
SHARD_HANDLE_BYTES = 256


class Stats(C.Structure):
    """``ic_stats`` of include/imageclust_b200.h."""
    _fields_ = [
        ("n_items", C.c_int64), ("dim", C.c_int64),
        ("n_target", C.c_int32), ("n_merges", C.c_int32), ("n_final", C.c_int32), ("n_out", C.c_int32),
        ("exhausted", C.c_int32), ("n_near_ties", C.c_int32), ("n_rescans", C.c_int32), ("gram_mode", C.c_int32),
        ("near_tie_tol", C.c_float),
        ("ms_h2d", C.c_float), ("ms_prep", C.c_float), ("ms_gram", C.c_float), ("ms_nn_init", C.c_float),
        ("ms_loop", C.c_float), ("ms_d2h", C.c_float), ("ms_host", C.c_float), ("ms_total", C.c_float),
        ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("matrix_bytes", C.c_int64), ("n_iterations", C.c_int32), ("loop_mode", C.c_int32),
        ("exact", C.c_int32), ("n_horizon_raises", C.c_int32), ("n_exact", C.c_int64),
        ("n_filter_viol", C.c_int32), ("n_order_viol", C.c_int32), ("n_cut", C.c_int32), ("filter_max_err", C.c_float),
        ("horizon", C.c_double), ("ms_refine", C.c_float), ("n_restarts", C.c_int32), ("n_compactions", C.c_int32), ("ms_compact", C.c_float), ("ms_loop_kernel", C.c_float), ("loop_launches", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class LibraryMissing(RuntimeError):
    pass


_lib = None


def load():
    """Load the C-ABI library; raises ``LibraryMissing`` if it has not been built
    (``python -m imageclust_b200.build``).  Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(f"{LIB_PATH} not built: run `python -m imageclust_b200.build` (needs nvcc)")
    L = C.CDLL(os.environ.get("IC_LIB_PATH", LIB_PATH))  # IC_LIB_PATH: tuning builds of the same library
    vp, i32p, i64p, fp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float)
    i64, i32 = C.c_int64, C.c_int
    sig = {
        "ic_create": (i32, [C.POINTER(vp), i32]),
        "ic_destroy": (None, [vp]),
        "ic_last_error": (C.c_char_p, [vp]),
        "ic_pinned_alloc": (vp, [C.c_size_t]),
        "ic_pinned_free": (None, [vp]),
        "ic_set_option": (i32, [vp, C.c_char_p, C.c_double]),
        "ic_optimal_clusters": (i32, [i64, i64, i64, i64p]),
        "ic_cluster_with_constraints": (i32, [vp, vp, i64, i64, i64, i64, i64, i32p, i32p, i32p, C.POINTER(Stats)]),
        "ic_load": (i32, [vp, vp, i64, i64, i64]),
        "ic_load_device": (i32, [vp, vp, i64, i64, i64]),
        "ic_initial_distances": (i32, [vp, i32, i64]),
        "ic_set_matrix": (i32, [vp, fp, i64]),
        "ic_nn_init": (i32, [vp]),
        "ic_find_closest": (i32, [vp, i32p, i32p, fp]),
        "ic_merge_loop": (i32, [vp, i64, i64, i64]),
        "ic_run_resident": (i32, [vp, i64, i64, i32p, i32p, i32p, C.POINTER(Stats)]),
        "ic_build_clusters": (i32, [vp, i64, i32p, i32p, i32p]),
        "ic_read_matrix": (i32, [vp, fp, i64]),
        "ic_read_slots": (i32, [vp, i32p, i32p]),
        "ic_get_merge_trace": (i32, [vp, i32p, i32p, fp, i32p, fp, i64, i64p]),
        "ic_get_linkage": (i32, [vp, C.POINTER(C.c_double), i64, i64p]),
        "ic_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "ic_get_loop_profile": (i32, [vp, i64p]),
        "ic_time_kernel": (i32, [vp, C.c_char_p, i32, fp]),
        "ic_shard_init": (i32, [vp, i32, i32]),
        "ic_shard_export": (i32, [vp, vp]),
        "ic_shard_connect": (i32, [vp, vp]),
        "ic_shard_rows": (i32, [vp, i64p, i64p]),
        "ic_load_combined": (i32, [vp, vp, i64, i64, i64, i32p, i32p, i64]),
        "ic_read_x": (i32, [vp, fp, i64]),
        "ic_get_loop_block_waits": (i32, [vp, i64p, i64, i64p]),
    }
    assert sorted(sig) == sorted(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
