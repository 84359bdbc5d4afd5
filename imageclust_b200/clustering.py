"""Host-side mirror of the reference's ``internal/clustering`` package.

Same two entry points, same argument meaning and error behaviour as
``/root/reference/internal/clustering/clustering.go``:

* :func:`calculate_optimal_clusters`  <-  ``CalculateOptimalClusters`` (clustering.go:168-186)
* :func:`perform_clustering_with_constraints`  <-  ``PerformClusteringWithConstraints``
  (clustering.go:198-284): ``(map[int][]string, bool)`` becomes ``(dict[int, list[str]] | None, bool)``.

All computation happens in the CUDA library behind the C ABI
(``include/imageclust_b200.h``); this module only flattens the rows into a pinned
staging buffer, makes ONE call, and turns two int arrays into the map -- exactly
what the Go/cgo shim does (``go/clustering_cgo.go``).  There is no
CPU fallback: without the built library or without a GPU these functions raise.
"""
from __future__ import annotations

import ctypes as C
import logging
import threading
from dataclasses import dataclass

import numpy as np

from . import _lib

log = logging.getLogger("imageclust_b200.clustering")

_ERR_TEXT = {
    _lib.IC_ERR_TOO_FEW: "total items less than minimum cluster size",
    _lib.IC_ERR_UNSAT: "cannot satisfy cluster size constraints",
    _lib.IC_ERR_BAD_ARG: "bad argument",
    _lib.IC_ERR_CUDA: "CUDA failure",
    _lib.IC_ERR_OOM: "problem does not fit the device",
    _lib.IC_ERR_STATE: "call out of order",
    _lib.IC_ERR_TIMEOUT: "device watchdog",
    _lib.IC_ERR_INTERNAL: "internal error",
}


class EngineError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"imageclust_b200 error {code} ({_ERR_TEXT.get(code, '?')}): {text}")
        self.code = code


class ConstraintError(ValueError):
    """CalculateOptimalClusters returned an error (clustering.go:169-177)."""

    def __init__(self, code: int, text: str):
        super().__init__(text)
        self.code = code


def calculate_optimal_clusters(total_items: int, min_size: int, max_size: int):
    """``CalculateOptimalClusters(totalItems, minSize, maxSize) (int, error)``, clustering.go:168-186.

    Returns ``(n_clusters, None)`` or ``(0, message)``.  Pure host arithmetic inside the C
    library (``ic_optimal_clusters``), no GPU needed.
    """
    out = C.c_int64(0)
    rc = _lib.load().ic_optimal_clusters(int(total_items), int(min_size), int(max_size), C.byref(out))
    if rc == _lib.IC_OK:
        return int(out.value), None
    if rc == _lib.IC_ERR_TOO_FEW:
        return 0, f"total items ({total_items}) less than minimum cluster size ({min_size})"
    if rc == _lib.IC_ERR_UNSAT:
        return 0, (f"cannot satisfy cluster size constraints with total items ({total_items}), "
                   f"minSize ({min_size}), and maxSize ({max_size})")
    return 0, f"invalid arguments: total items ({total_items}), minSize ({min_size}), maxSize ({max_size})"


@dataclass
class MergeTrace:
    key_hi: np.ndarray
    key_lo: np.ndarray
    dist: np.ndarray
    size: np.ndarray
    gap: np.ndarray


@dataclass
class ClusterResult:
    clusters: list  # list of int32 arrays of item indices; map id == list position
    stats: dict


class Engine:
    """One ``ic_ctx``: a device, a stream and the resident problem.  Not thread safe;
    :func:`perform_clustering_with_constraints` serialises calls on a shared engine."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.ic_create(C.byref(h), int(device))
        if rc != _lib.IC_OK or not h:
            raise EngineError(rc, "ic_create failed: a CUDA device of compute capability 10.x is required "
                                  "(there is no CPU fallback)")
        self._h = h
        self._pinned = []
        self.n = 0
        self.d = 0

    # -- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.ic_destroy(self._h)
            self._h = None
        for p in getattr(self, "_pinned", []):
            self._L.ic_pinned_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != _lib.IC_OK:
            raise EngineError(rc, (self._L.ic_last_error(self._h) or b"").decode())

    def set_option(self, name: str, value: float):
        self._check(self._L.ic_set_option(self._h, name.encode(), float(value)))

    def pinned_empty(self, shape) -> np.ndarray:
        """fp32 array in page-locked host memory (the H2D staging buffer of the shim)."""
        count = int(np.prod(shape))
        p = self._L.ic_pinned_alloc(max(count, 1) * 4)
        if not p:
            raise MemoryError("ic_pinned_alloc failed")
        self._pinned.append(p)
        buf = (C.c_float * max(count, 1)).from_address(p)
        return np.frombuffer(buf, dtype=np.float32, count=count).reshape(shape)

    # -- the whole path -------------------------------------------------------
    def cluster(self, x: np.ndarray, min_size: int, max_size: int) -> ClusterResult:
        """``ic_cluster_with_constraints`` on a host matrix (pinned or pageable)."""
        x = self._as_matrix(x)
        n, d = x.shape
        offsets = np.zeros(n + 1, np.int32)
        members = np.zeros(max(n, 1), np.int32)
        k = C.c_int32(0)
        st = _lib.Stats()
        rc = self._L.ic_cluster_with_constraints(self._h, x.ctypes.data_as(C.c_void_p), n, d, x.strides[0] // 4,
                                                 int(min_size), int(max_size), _i32p(offsets), _i32p(members),
                                                 C.byref(k), C.byref(st))
        if rc in (_lib.IC_ERR_TOO_FEW, _lib.IC_ERR_UNSAT):
            raise ConstraintError(rc, calculate_optimal_clusters(n, min_size, max_size)[1])
        self._check(rc)
        self.n, self.d = n, d
        return ClusterResult(_split(offsets, members, k.value), st.as_dict())

    def run_resident(self, min_size: int, max_size: int) -> ClusterResult:
        """The path on the matrix already in HBM (after :meth:`load` / :meth:`load_device`)."""
        n = self.n
        offsets = np.zeros(n + 1, np.int32)
        members = np.zeros(max(n, 1), np.int32)
        k = C.c_int32(0)
        st = _lib.Stats()
        rc = self._L.ic_run_resident(self._h, int(min_size), int(max_size), _i32p(offsets), _i32p(members),
                                     C.byref(k), C.byref(st))
        if rc in (_lib.IC_ERR_TOO_FEW, _lib.IC_ERR_UNSAT):
            raise ConstraintError(rc, calculate_optimal_clusters(n, min_size, max_size)[1])
        self._check(rc)
        return ClusterResult(_split(offsets, members, k.value), st.as_dict())

    # -- staged entry points (unit parity with the reference's functions) ----------
    @staticmethod
    def _as_matrix(x) -> np.ndarray:
        x = np.asarray(x)
        if x.ndim != 2:
            raise ValueError("embeddings must be a [N x D] matrix")
        if x.dtype != np.float32 or x.strides[1] != 4 or x.strides[0] % 4 or x.strides[0] < 4 * x.shape[1]:
            x = np.ascontiguousarray(x, np.float32)
        return x

    def load(self, x):
        x = self._as_matrix(x)
        self._check(self._L.ic_load(self._h, x.ctypes.data_as(C.c_void_p), x.shape[0], x.shape[1],
                                    x.strides[0] // 4))
        self.n, self.d = x.shape

    def load_combined(self, image_embeddings, item_labels, label_set):
        """``GenerateLabelVector`` + ``CombineEmbeddings`` (embeddings.go:166-183, workflow.go:167-168) on the
        device: ``image_embeddings`` is ``[N x D_img]``, ``item_labels[i]`` the label names of item ``i``,
        ``label_set`` the ``name -> index`` map (``BuildLabelSet``).  Only the image block and the label indices
        are uploaded; the ``[N x (D_img + len(label_set))]`` matrix is formed in HBM."""
        img = self._as_matrix(image_embeddings)
        n, d_img = img.shape
        if len(item_labels) != n:
            raise ValueError("one label list per item")
        offsets = np.zeros(n + 1, np.int32)
        ids = []
        for i, labels in enumerate(item_labels):
            # the map lookup of embeddings.go:169; a label outside the set becomes -1 and is ignored by the kernel
            ids.extend(label_set.get(name, -1) for name in labels)
            offsets[i + 1] = len(ids)
        ids = np.asarray(ids if ids else [0], np.int32)
        self._check(self._L.ic_load_combined(self._h, img.ctypes.data_as(C.c_void_p), n, d_img, img.strides[0] // 4,
                                             _i32p(offsets), _i32p(ids), len(label_set)))
        self.n, self.d = n, d_img + len(label_set)

    def read_x(self) -> np.ndarray:
        out = np.zeros((self.n, max(self.d, 1)), np.float32)
        self._check(self._L.ic_read_x(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), max(self.d, 1)))
        return out[:, :self.d]

    def load_device(self, ptr: int, n: int, d: int, ldx: int):
        self._check(self._L.ic_load_device(self._h, C.c_void_p(ptr), n, d, ldx))
        self.n, self.d = n, d

    def initial_distances(self, mode: int = _lib.GRAM_TCGEN05_3XTF32, max_size: int = 1 << 30):
        """ComputeInitialDistanceMatrix (clustering.go:61-73) into HBM."""
        self._check(self._L.ic_initial_distances(self._h, int(mode), int(max_size)))

    def set_matrix(self, m):
        m = np.ascontiguousarray(m, np.float32)
        assert m.shape == (self.n, self.n)
        self._check(self._L.ic_set_matrix(self._h, m.ctypes.data_as(C.POINTER(C.c_float)), self.n))

    def nn_init(self):
        self._check(self._L.ic_nn_init(self._h))

    def find_closest(self):
        """FindClosestClusters (clustering.go:119-133) -> (key_hi, key_lo, dist); key_hi == -1 if none."""
        hi, lo, d = C.c_int32(-1), C.c_int32(-1), C.c_float(0)
        self._check(self._L.ic_find_closest(self._h, C.byref(hi), C.byref(lo), C.byref(d)))
        return hi.value, lo.value, d.value

    def merge_loop(self, min_size: int, max_size: int, max_merges: int = -1):
        rc = self._L.ic_merge_loop(self._h, int(min_size), int(max_size), int(max_merges))
        if rc in (_lib.IC_ERR_TOO_FEW, _lib.IC_ERR_UNSAT):
            raise ConstraintError(rc, calculate_optimal_clusters(self.n, min_size, max_size)[1])
        self._check(rc)

    def build_clusters(self, min_size: int):
        offsets = np.zeros(self.n + 1, np.int32)
        members = np.zeros(max(self.n, 1), np.int32)
        k = C.c_int32(0)
        self._check(self._L.ic_build_clusters(self._h, int(min_size), _i32p(offsets), _i32p(members), C.byref(k)))
        return _split(offsets, members, k.value)

    def read_matrix(self) -> np.ndarray:
        out = np.zeros((self.n, self.n), np.float32)
        self._check(self._L.ic_read_matrix(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), self.n))
        return out

    def read_slots(self):
        key = np.zeros(max(self.n, 1), np.int32)
        size = np.zeros(max(self.n, 1), np.int32)
        self._check(self._L.ic_read_slots(self._h, _i32p(key), _i32p(size)))
        return key[:self.n], size[:self.n]

    def merge_trace(self) -> MergeTrace:
        cap = max(self.n, 1)
        hi, lo, sz = (np.zeros(cap, np.int32) for _ in range(3))
        dist, gap = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        m = C.c_int64(0)
        self._check(self._L.ic_get_merge_trace(self._h, _i32p(hi), _i32p(lo), dist.ctypes.data_as(C.POINTER(C.c_float)),
                                               _i32p(sz), gap.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(m)))
        k = m.value
        return MergeTrace(hi[:k], lo[:k], dist[:k], sz[:k], gap[:k])

    def linkage(self) -> np.ndarray:
        """The merge trace as a scipy-style linkage matrix ``[n_merges x 4]`` (``ic_get_linkage``): rows
        ``{key_lo, key_hi, sqrt(2 d), size}``; an unconstrained run equals ``scipy.cluster.hierarchy.linkage(X, "ward")``."""
        cap = max(self.n, 1)
        z = np.zeros((cap, 4), np.float64)
        m = C.c_int64(0)
        self._check(self._L.ic_get_linkage(self._h, z.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(m)))
        return z[:int(m.value)].copy()

    def stats(self) -> dict:
        st = _lib.Stats()
        self._check(self._L.ic_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def loop_profile(self) -> dict:
        """Cycles block 0 spent per phase of the last merge-loop launch (option profile_loop=1)."""
        out = (C.c_int64 * 16)()
        self._check(self._L.ic_get_loop_profile(self._h, out))
        # batched loop (merge_batch.cu): publish = phase 1 (new rows' lists + rescans), exchange = phase 2 (heads), update = batch
        # selection, scan = phase 3 + 4 (rows .. end of iteration); pub_argmin = Lance-Williams rows + centroids, pub_reduce /
        # pub_fence = barrier waits after phase 1 / 2, pub_stores = end of validation .. end of iteration, of which exch_poll =
        # wait for the slowest block's rows phase and exch_rank = this block's share of the exact phase
        return dict(zip(("publish", "exchange", "update", "scan", "fold", "merges", "iterations", "rescans",
                         "reserved", "bubbles", "pub_argmin", "pub_reduce", "pub_fence", "pub_stores", "exch_poll",
                         "exch_rank"), list(out)))

    # -- row-block sharding over several GPUs: one process (and Engine) per GPU ------------------
    def shard_init(self, rank: int, world: int):
        """Declare this engine rank ``rank`` of ``world`` row-block shards (before :meth:`load`)."""
        self._check(self._L.ic_shard_init(self._h, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def shard_export(self) -> bytes:
        """CUDA-IPC handles of this rank's row block and mailbox (after :meth:`load`)."""
        buf = C.create_string_buffer(_lib.SHARD_HANDLE_BYTES)
        self._check(self._L.ic_shard_export(self._h, C.cast(buf, C.c_void_p)))
        return buf.raw

    def shard_connect(self, blobs):
        """Peer-map the other ranks' row blocks; ``blobs`` = every rank's export, in rank order."""
        raw = b"".join(blobs)
        assert len(raw) == self.world * _lib.SHARD_HANDLE_BYTES
        buf = C.create_string_buffer(raw, len(raw))
        self._check(self._L.ic_shard_connect(self._h, C.cast(buf, C.c_void_p)))

    def shard_rows(self):
        lo, hi = C.c_int64(0), C.c_int64(0)
        self._check(self._L.ic_shard_rows(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def loop_block_waits(self) -> np.ndarray:
        """Cycles every merge-loop block waited in the exchanges of the last launch (option profile_loop=1)."""
        out = (C.c_int64 * 256)()
        nb = C.c_int64(0)
        self._check(self._L.ic_get_loop_block_waits(self._h, out, 256, C.byref(nb)))
        return np.array(list(out)[:min(nb.value, 240)], np.int64)

    def time_kernel(self, which: str, repeats: int = 1) -> float:
        ms = C.c_float(0)
        self._check(self._L.ic_time_kernel(self._h, which.encode(), int(repeats), C.byref(ms)))
        return ms.value


def trace_to_linkage(trace) -> np.ndarray:
    """Host-side twin of ``ic_get_linkage`` for any merge trace with ``key_hi / key_lo / dist / size`` arrays (the
    device's :class:`MergeTrace` or an oracle result)."""
    m = len(trace.key_hi)
    z = np.zeros((m, 4), np.float64)
    z[:, 0] = np.asarray(trace.key_lo[:m], np.float64)
    z[:, 1] = np.asarray(trace.key_hi[:m], np.float64)
    z[:, 2] = np.sqrt(2.0 * np.asarray(trace.dist[:m], np.float64))
    z[:, 3] = np.asarray(trace.size[:m], np.float64)
    return z


def clustering_report(n_items: int, clusters, stats: dict | None = None) -> dict:
    """What the reference's caller cannot see today (workflow.go:89-97): the items that vanished because
    their cluster stayed below ``minSize`` (clustering.go:268-271), and the near-tie count of the run."""
    seen = np.zeros(int(n_items), bool)
    for c in clusters:
        seen[np.asarray(c, np.int64)] = True
    rep = {"n_items": int(n_items), "n_clusters": len(clusters), "dropped_items": np.flatnonzero(~seen).astype(np.int32)}
    if stats is not None:
        rep.update(n_near_ties=stats.get("n_near_ties"), near_tie_tol=stats.get("near_tie_tol"),
                   exhausted=bool(stats.get("exhausted")), n_final=stats.get("n_final"))
    return rep


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _split(offsets, members, k):
    return [members[offsets[i]:offsets[i + 1]].copy() for i in range(k)]


_shared = None
_shared_lock = threading.Lock()


def _shared_engine() -> Engine:
    global _shared
    if _shared is None:
        _shared = Engine(0)
    return _shared


def perform_clustering_with_constraints(embeddings, product_reference_ids, min_size: int, max_size: int,
                                        engine: Engine | None = None):
    """``PerformClusteringWithConstraints(embeddings, productReferenceIDs, minSize, maxSize)``
    (clustering.go:198-284) -> ``(cluster_map, ok)``.

    ``cluster_map[id]`` lists the reference ids of cluster ``id`` (ids dense from 0 in the
    reference's slice order, members in the reference's order; items of clusters below
    ``min_size`` are absent, clustering.go:268-271).  Returns ``(None, False)`` where the
    reference returns ``(nil, false)`` (constraint errors, :204-207), and also -- instead of
    the Go runtime panic -- for ragged rows (:149-151) or fewer ids than rows (:276).
    Device failures raise :class:`EngineError` (there is no CPU path to fall back to).
    """
    try:
        rows = len(embeddings)
    except TypeError:
        return None, False
    log.info("Total items for clustering: %d", rows)
    n_clusters, err = calculate_optimal_clusters(rows, min_size, max_size)
    if err is not None:
        log.warning("Clustering constraint error: %s", err)
        return None, False
    log.info("Optimal number of clusters calculated: %d", n_clusters)
    if len(product_reference_ids) < rows:
        log.error("productReferenceIDs shorter than embeddings (%d < %d)", len(product_reference_ids), rows)
        return None, False
    widths = {len(r) for r in embeddings} if not isinstance(embeddings, np.ndarray) else {embeddings.shape[1]}
    if len(widths) != 1:
        log.error("embeddings have different lengths: %s", sorted(widths)[:4])
        return None, False
    d = widths.pop()
    with _shared_lock:
        eng = engine or _shared_engine()
        # flatten the rows into the pinned staging buffer (the cgo shim does the same: Go
        # pointers never cross the boundary)
        stage = eng.pinned_empty((rows, d))
        try:
            if isinstance(embeddings, np.ndarray):
                np.copyto(stage, embeddings, casting="same_kind")
            else:
                for i, r in enumerate(embeddings):
                    stage[i] = r
            res = eng.cluster(stage, min_size, max_size)
        except ConstraintError as e:  # unreachable: checked above
            log.warning("Clustering constraint error: %s", e)
            return None, False
        finally:
            eng._L.ic_pinned_free(eng._pinned.pop())
    cluster_map = {cid: [product_reference_ids[i] for i in idx] for cid, idx in enumerate(res.clusters)}
    log.info("Clustering successful. Formed %d valid clusters.", len(cluster_map))
    return cluster_map, True
