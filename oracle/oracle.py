"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  It wraps ``oracle/liboracle.so`` (built from
``ward_literal.c`` + ``ward_fast.c`` by ``oracle/Makefile``), the C restatement of
``/root/reference/internal/clustering/clustering.go``.

PARITY UNPINNED by the reference's own tests (it has none);
the restatement is pinned by tests/test_oracle_*.py instead.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ERR_TOO_FEW, ERR_UNSAT, ERR_BAD_ARG = -1, -2, -3
FAST_EAGER, FAST_LW, FAST_LW32 = 1, 2, 4


class _Stats(C.Structure):
    _fields_ = [("n_target", C.c_int), ("n_merges", C.c_int), ("n_rejections", C.c_long),
                ("exhausted", C.c_int), ("n_final", C.c_int), ("n_out", C.c_int)]


class _DeviceStats(C.Structure):
    _fields_ = [("n_iterations", C.c_long), ("n_exact", C.c_long), ("n_raises", C.c_long), ("n_cut", C.c_long),
                ("n_violations", C.c_long), ("max_filter_err", C.c_double), ("horizon", C.c_double)]


class _Trace(C.Structure):
    _fields_ = [("key_hi", C.POINTER(C.c_int)), ("key_lo", C.POINTER(C.c_int)),
                ("pos_i", C.POINTER(C.c_int)), ("pos_j", C.POINTER(C.c_int)),
                ("dist", C.POINTER(C.c_float)), ("size", C.POINTER(C.c_int)),
                ("gap", C.POINTER(C.c_float))]


def build(force: bool = False) -> str:
    """Compile liboracle.so if missing or stale. Returns its path."""
    srcs = [os.path.join(_HERE, f) for f in ("ward_literal.c", "ward_fast.c", "ward_device.c", "oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        L.oracle_dot_f32.restype = C.c_float
        L.oracle_dot_f32.argtypes = [fp, fp, C.c_int]
        L.oracle_ward_distance.restype = C.c_float
        L.oracle_ward_distance.argtypes = [fp, C.c_long, fp, C.c_long, C.c_int]
        L.oracle_merge_centroid.restype = None
        L.oracle_merge_centroid.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, fp]
        L.oracle_optimal_clusters.restype = C.c_int
        L.oracle_optimal_clusters.argtypes = [C.c_long, C.c_long, C.c_long, C.POINTER(C.c_long)]
        L.oracle_literal_cluster.restype = C.c_int
        L.oracle_literal_cluster.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, ip,
                                             C.POINTER(_Trace), fp, fp, ip, C.POINTER(_Stats)]
        L.oracle_fast_cluster_ex.restype = C.c_int
        L.oracle_fast_cluster_ex.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp,
                                             ip, ip, ip, C.POINTER(_Trace), C.POINTER(_Stats)]
        L.oracle_device_cluster.restype = C.c_int
        L.oracle_device_cluster.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_double, C.c_double,
                                            C.c_double, C.c_int, C.c_int, ip, ip, ip, C.POINTER(_Trace),
                                            C.POINTER(_Stats), C.POINTER(_DeviceStats)]
        L.oracle_initial_matrix.restype = C.c_int
        L.oracle_initial_matrix.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


@dataclass
class OracleResult:
    ok: bool
    rc: int
    clusters: list = field(default_factory=list)  # list of int32 arrays (item indices), map id = position
    key_hi: np.ndarray | None = None
    key_lo: np.ndarray | None = None
    pos_i: np.ndarray | None = None
    pos_j: np.ndarray | None = None
    dist: np.ndarray | None = None
    size: np.ndarray | None = None
    gap: np.ndarray | None = None
    n_target: int = 0
    n_merges: int = 0
    n_rejections: int = 0
    exhausted: bool = False
    n_final: int = 0
    init_matrix: np.ndarray | None = None
    final_matrix: np.ndarray | None = None
    final_keys: np.ndarray | None = None


def optimal_clusters(total: int, min_size: int, max_size: int):
    """CalculateOptimalClusters (clustering.go:168-186) -> (n, rc)."""
    out = C.c_long(0)
    rc = lib().oracle_optimal_clusters(total, min_size, max_size, C.byref(out))
    return (int(out.value) if rc == 0 else 0), rc


def dot_f32(a, b) -> float:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().oracle_dot_f32(_fp(a), _fp(b), a.size))


def ward_distance(ca, size_a, cb, size_b) -> float:
    ca = np.ascontiguousarray(ca, np.float32)
    cb = np.ascontiguousarray(cb, np.float32)
    return float(lib().oracle_ward_distance(_fp(ca), size_a, _fp(cb), size_b, ca.size))


def merge_centroid(ca, size_a, cb, size_b):
    ca = np.ascontiguousarray(ca, np.float32)
    cb = np.ascontiguousarray(cb, np.float32)
    out = np.empty_like(ca)
    lib().oracle_merge_centroid(_fp(ca), size_a, _fp(cb), size_b, ca.size, _fp(out))
    return out


def initial_matrix(x, n_threads: int = 0):
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    out = np.zeros((n, n), np.float32)
    lib().oracle_initial_matrix(_fp(x), n, d, n_threads or (os.cpu_count() or 1), _fp(out))
    return out


def _alloc_trace(n, with_gap):
    arrs = dict(key_hi=np.full(n, -1, np.int32), key_lo=np.full(n, -1, np.int32),
                pos_i=np.full(n, -1, np.int32), pos_j=np.full(n, -1, np.int32),
                dist=np.zeros(n, np.float32), size=np.zeros(n, np.int32),
                gap=np.zeros(n, np.float32) if with_gap else None)
    tr = _Trace(_ip(arrs["key_hi"]), _ip(arrs["key_lo"]), _ip(arrs["pos_i"]), _ip(arrs["pos_j"]),
                _fp(arrs["dist"]), _ip(arrs["size"]), _fp(arrs["gap"]) if with_gap else None)
    return arrs, tr


def _finish(rc, st, arrs, offsets, members, n_out, **extra):
    if rc != 0:
        return OracleResult(ok=False, rc=rc)
    k = n_out.value
    clusters = [members[offsets[i]:offsets[i + 1]].copy() for i in range(k)]
    m = st.n_merges
    return OracleResult(ok=True, rc=0, clusters=clusters,
                        key_hi=arrs["key_hi"][:m], key_lo=arrs["key_lo"][:m],
                        pos_i=arrs["pos_i"][:m], pos_j=arrs["pos_j"][:m],
                        dist=arrs["dist"][:m], size=arrs["size"][:m],
                        gap=None if arrs["gap"] is None else arrs["gap"][:m],
                        n_target=st.n_target, n_merges=m, n_rejections=st.n_rejections,
                        exhausted=bool(st.exhausted), n_final=st.n_final, **extra)


def literal_cluster(x, min_size: int, max_size: int, want_matrices: bool = False) -> OracleResult:
    """PerformClusteringWithConstraints (clustering.go:198-284), literal restatement."""
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    offsets = np.zeros(n + 1, np.int32)
    members = np.zeros(max(n, 1), np.int32)
    n_out = C.c_int(0)
    st = _Stats()
    arrs, tr = _alloc_trace(max(n, 1), False)
    init_m = np.zeros((n, n), np.float32) if want_matrices else None
    fin_m = np.zeros((n, n), np.float32) if want_matrices else None
    fin_k = np.zeros(max(n, 1), np.int32) if want_matrices else None
    rc = lib().oracle_literal_cluster(_fp(x), n, d, min_size, max_size, _ip(offsets), _ip(members),
                                      C.byref(n_out), C.byref(tr), _fp(init_m), _fp(fin_m), _ip(fin_k),
                                      C.byref(st))
    extra = {}
    if want_matrices and rc == 0:
        nf = st.n_final
        extra = dict(init_matrix=init_m, final_matrix=fin_m.reshape(-1)[:nf * nf].reshape(nf, nf).copy(),
                     final_keys=fin_k[:nf].copy())
    return _finish(rc, st, arrs, offsets, members, n_out, **extra)


def fast_cluster(x, min_size: int, max_size: int, flags: int = 0, n_threads: int = 0,
                 init_matrix=None) -> OracleResult:
    """Same semantics as literal_cluster with an NN cache (ward_fast.c)."""
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    offsets = np.zeros(n + 1, np.int32)
    members = np.zeros(max(n, 1), np.int32)
    n_out = C.c_int(0)
    st = _Stats()
    arrs, tr = _alloc_trace(max(n, 1), True)
    if init_matrix is not None:
        init_matrix = np.ascontiguousarray(init_matrix, np.float32)
        assert init_matrix.shape == (n, n)
    rc = lib().oracle_fast_cluster_ex(_fp(x), n, d, min_size, max_size, flags,
                                      n_threads or (os.cpu_count() or 1), _fp(init_matrix),
                                      _ip(offsets), _ip(members), C.byref(n_out), C.byref(tr), C.byref(st))
    return _finish(rc, st, arrs, offsets, members, n_out)


def device_cluster(x, min_size: int, max_size: int, init_matrix=None, horizon_factor: float = 1.25,
                   eps_filter: float = 3e-5, delta_cut: float = 8e-6, max_batch: int = 512, n_threads: int = 0):
    """CPU restatement of the device's merge loop (ward_device.c) -> (OracleResult, device-stats dict)."""
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    offsets = np.zeros(n + 1, np.int32)
    members = np.zeros(max(n, 1), np.int32)
    n_out = C.c_int(0)
    st = _Stats()
    ds = _DeviceStats()
    arrs, tr = _alloc_trace(max(n, 1), True)
    if init_matrix is not None:
        init_matrix = np.ascontiguousarray(init_matrix, np.float32)
        assert init_matrix.shape == (n, n)
    rc = lib().oracle_device_cluster(_fp(x), n, d, min_size, max_size, _fp(init_matrix), horizon_factor, eps_filter,
                                     delta_cut, max_batch, n_threads or (os.cpu_count() or 1), _ip(offsets),
                                     _ip(members), C.byref(n_out), C.byref(tr), C.byref(st), C.byref(ds))
    return _finish(rc, st, arrs, offsets, members, n_out), {f: getattr(ds, f) for f, _ in _DeviceStats._fields_}


# ---- input formation: literal restatement, plain Python loops (tiny inputs only) ----------------

def generate_label_vector(labels, label_set):
    """``GenerateLabelVector`` -- /root/reference/internal/embeddings/embeddings.go:166-174."""
    label_vector = [np.float32(0.0)] * len(label_set)       # make([]float32, len(labelSet))          :167
    for label in labels:                                    # for _, label := range labels            :168
        if label in label_set:                              # if idx, exists := labelSet[label]       :169
            label_vector[label_set[label]] = np.float32(1.0)  # labelVector[idx] = 1.0                :170
    return np.asarray(label_vector, np.float32)


def combine_embeddings(embedding, label_vector):
    """``CombineEmbeddings`` -- /root/reference/internal/embeddings/embeddings.go:177-183."""
    combined = np.zeros(len(embedding) + len(label_vector), np.float32)  # make(..., len(e)+len(l))  :179
    combined[:len(embedding)] = embedding                                 # copy(combined, embedding) :180
    combined[len(embedding):] = label_vector                              # copy(combined[len(e):], l) :181
    return combined
