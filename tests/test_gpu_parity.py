"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle
and the committed golden vectors.

The product path (computed initial matrix, batched loop, option "exact" = 1, the default) must reproduce the oracle in
REFERENCE ARITHMETIC -- ``oracle.fast_cluster(flags=0)``, bit-identical to the literal restatement of clustering.go --
merge for merge, distance bits included.  The Lance-Williams machinery underneath (supplied matrices, "exact" = 0, the
one-merge-per-iteration loop, virtual ranks) is checked against the oracle's Lance-Williams mode, which restates the
device's arithmetic.  <= 1e-5 relative (BASELINE north_star) for the tensor-core Gram kernels."""
import contextlib

import numpy as np
import pytest

from imageclust_b200 import _lib, clustering, synth
from tests.helpers import ari, golden_clusters, golden_names, load_golden, same_clusters

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: "<= 1e-5 rel, fp32"
SMALL_GOLDENS = [n for n in golden_names() if n != "cfgA_1000x2048"]
LW_EAGER = 3  # oracle.FAST_EAGER | oracle.FAST_LW


@pytest.fixture(scope="module")
def eng():
    e = clustering.Engine(0)
    yield e
    e.close()


@contextlib.contextmanager
def lw_only(e):
    """Lance-Williams values only (no horizon, no centroid re-evaluation): the arithmetic the LW oracle restates."""
    e.set_option("exact", 0)
    try:
        yield
    finally:
        e.set_option("exact", 1)


def _assert_reference_run(st):
    """The horizon's guarantees held: no stored value was off by more than the filter tolerance, no pair created
    inside a batch preceded a later member of it."""
    assert st["exact"] == 1 and st["n_filter_viol"] == 0 and st["n_order_viol"] == 0, st


def _same_trace(tr, o, exact_dist=True):
    assert len(tr.key_hi) == o.n_merges
    assert np.array_equal(tr.key_hi, o.key_hi), np.flatnonzero(tr.key_hi != o.key_hi)[:5]
    assert np.array_equal(tr.key_lo, o.key_lo)
    assert np.array_equal(tr.size, o.size)
    if exact_dist:
        assert np.array_equal(tr.dist.view(np.uint32), o.dist.view(np.uint32))
    else:
        np.testing.assert_allclose(tr.dist, o.dist, rtol=RTOL)


# ---- K1: ComputeInitialDistanceMatrix (clustering.go:61-73) -----------------------------

@pytest.mark.parametrize("name", SMALL_GOLDENS)
def test_exact_gram_is_bit_identical_to_the_reference_arithmetic(eng, name):
    g = load_golden(name)
    eng.load(g["x"])
    eng.initial_distances(_lib.GRAM_EXACT_FP32)
    m = eng.read_matrix()
    assert np.array_equal(m.view(np.uint32), g["init_matrix"].view(np.uint32))


def _gram_case(eng, oracle, x):
    eng.load(x)
    eng.initial_distances(_lib.GRAM_TCGEN05_3XTF32)
    m = eng.read_matrix()
    ref = oracle.initial_matrix(x)
    assert np.array_equal(m, m.T)
    assert np.all(np.diag(m) == 0)
    off = ~np.eye(len(x), dtype=bool)
    rel = np.abs(m[off] - ref[off]) / ref[off]
    return float(rel.max())


@pytest.mark.parametrize("n,d,relu", [(256, 64, False), (700, 2048, False), (700, 2048, True), (333, 2148, False),
                                      (1000, 96, False), (129, 40, False)])
def test_tcgen05_gram_within_tolerance(eng, oracle, n, d, relu):
    x = synth.gaussian_mixture(n, d, 5, 20, seed=100 + n + d, relu_like=relu)
    assert _gram_case(eng, oracle, x) <= RTOL


def test_tcgen05_gram_config_e_like(eng, oracle):
    x = synth.combined_features(600, 2048, 100, 2, 8, seed=7)
    assert _gram_case(eng, oracle, x) <= RTOL


def test_exact_gram_larger_than_one_tile(eng, oracle):
    x = synth.gaussian_mixture(300, 77, 3, 9, seed=9)
    eng.load(x)
    eng.initial_distances(_lib.GRAM_EXACT_FP32)
    assert np.array_equal(eng.read_matrix(), oracle.initial_matrix(x))


# ---- K2: FindClosestClusters (clustering.go:119-133) ------------------------------------

@pytest.mark.parametrize("name", SMALL_GOLDENS)
def test_find_closest_matches_first_merge(eng, name):
    g = load_golden(name)
    eng.load(g["x"])
    eng.set_matrix(g["init_matrix"])
    eng.nn_init()
    hi, lo, d = eng.find_closest()
    # the literal scan returns positions; before the first merge position == key
    if int(g["max_size"]) >= 2 and len(g["key_hi"]):
        m = g["init_matrix"]
        n = len(m)
        best = (np.float32(np.finfo(np.float32).max), -1, -1)
        for i in range(n):
            for j in range(i):
                if m[i, j] < best[0]:
                    best = (m[i, j], i, j)
        assert (hi, lo) == (best[1], best[2]) and np.float32(d) == best[0]


def test_find_closest_none_selectable(eng):
    x = synth.gaussian_mixture(40, 8, 1, 4, seed=1)
    eng.load(x)
    m = np.full((40, 40), np.finfo(np.float32).max, np.float32)  # MaxFloat32 never wins (clustering.go:120,124)
    m[3, 1] = m[1, 3] = np.inf
    m[5, 2] = m[2, 5] = np.nan
    eng.set_matrix(m)
    eng.nn_init()
    assert eng.find_closest()[0] == -1


# ---- K3: the merge loop (clustering.go:220-246) -----------------------------------------

@pytest.mark.parametrize("name", SMALL_GOLDENS)
def test_merge_loop_bit_exact_vs_oracle_lw(eng, oracle, name):
    """Same initial matrix in, the device loop and the oracle's Lance-Williams mode
    must produce the identical merge sequence, distances and cluster lists."""
    g = load_golden(name)
    mn, mx = int(g["min_size"]), int(g["max_size"])
    o = oracle.fast_cluster(g["x"], mn, mx, flags=LW_EAGER, init_matrix=g["init_matrix"])
    eng.load(g["x"])
    eng.set_matrix(g["init_matrix"])
    eng.nn_init()
    eng.merge_loop(mn, mx)
    tr = eng.merge_trace()
    _same_trace(tr, o)
    st = eng.stats()
    assert st["n_merges"] == o.n_merges and st["n_final"] == o.n_final and bool(st["exhausted"]) == o.exhausted
    cl = eng.build_clusters(mn)
    assert same_clusters(cl, o.clusters)
    assert st["exact"] == 0  # a supplied matrix need not belong to the resident X: Lance-Williams values only


@pytest.mark.parametrize("name", SMALL_GOLDENS)
@pytest.mark.parametrize("gram", [_lib.GRAM_EXACT_FP32, _lib.GRAM_TCGEN05_I8])
def test_full_path_equals_golden(eng, oracle, name, gram):
    """ic_cluster_with_constraints == the literal restatement of clustering.go (the committed goldens): merge sequence,
    distance bits, cluster lists, surviving keys -- unconditionally, with the exact-arithmetic Gram kernel and with the
    tensor-core one (whose values the horizon sweep replaces where they can matter)."""
    g = load_golden(name)
    mn, mx = int(g["min_size"]), int(g["max_size"])
    eng.set_option("gram_mode", gram)
    try:
        res = eng.cluster(g["x"], mn, mx)
    finally:
        eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)
    tr = eng.merge_trace()
    assert np.array_equal(tr.key_hi, g["key_hi"]) and np.array_equal(tr.key_lo, g["key_lo"])
    assert np.array_equal(tr.dist.view(np.uint32), g["dist"].view(np.uint32))
    assert same_clusters(res.clusters, golden_clusters(g))
    key, size = eng.read_slots()
    assert np.array_equal(np.sort(key[key >= 0]), np.sort(g["final_keys"]))
    assert res.stats["n_target"] == int(g["n_target"])
    if mx >= 2 and len(tr.key_hi):
        _assert_reference_run(res.stats)


@pytest.mark.parametrize("n,d,mn,mx,seed,gram", [
    (3000, 64, 4, 12, 1, _lib.GRAM_TCGEN05_I8), (2000, 2048, 10, 50, 2, _lib.GRAM_TCGEN05_I8),
    (2500, 100, 1, 2500, 3, _lib.GRAM_EXACT_FP32), (5000, 32, 6, 8, 4, _lib.GRAM_TCGEN05_3XTF32),
    (1500, 2148, 2, 8, 5, _lib.GRAM_TCGEN05_I8)])
def test_full_path_equals_reference_arithmetic(eng, oracle, n, d, mn, mx, seed, gram):
    """Mixtures, one unconstrained run, an exhaustion run: the device's merge trace IS the reference-arithmetic one."""
    x = synth.gaussian_mixture(n, d, mn, min(mx, 40), seed=seed)
    o = oracle.fast_cluster(x, mn, mx, flags=0)
    eng.set_option("gram_mode", gram)
    try:
        res = eng.cluster(x, mn, mx)
    finally:
        eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(res.clusters, o.clusters)
    assert bool(res.stats["exhausted"]) == o.exhausted
    _assert_reference_run(res.stats)


def test_staged_resume_equals_single_run(eng, oracle):
    x = synth.gaussian_mixture(500, 48, 3, 10, seed=33)
    o = oracle.fast_cluster(x, 3, 10, flags=0)
    eng.load(x)
    eng.initial_distances(_lib.GRAM_EXACT_FP32, 10)
    eng.nn_init()
    for step in (1, 7, 100, 0, 13):
        eng.merge_loop(3, 10, step)
    eng.merge_loop(3, 10)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(eng.build_clusters(3), o.clusters)


def test_many_rows_lose_their_partner(eng, oracle):
    """Duplicates of one point: every copy's nearest partner is the lowest-key copy, so its
    merge sends dozens of rows to the block-per-row rescan path."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((400, 8)).astype(np.float32)
    x[50:120] = x[7]
    x[200:230] = x[9]
    lit = oracle.literal_cluster(x, 1, 6)
    for gram in (_lib.GRAM_EXACT_FP32, _lib.GRAM_TCGEN05_I8):  # the tensor-core Gram does not give duplicates exactly 0
        eng.set_option("gram_mode", gram)
        try:
            res = eng.cluster(x, 1, 6)
        finally:
            eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)
        _same_trace(eng.merge_trace(), lit)
        assert same_clusters(res.clusters, lit.clusters)


@pytest.mark.parametrize("n,d,mn,mx,threads", [(3000, 64, 4, 12, 0), (3000, 64, 4, 12, 512), (5000, 32, 6, 8, 0),
                                               (2500, 100, 1, 2500, 0)])
def test_loop_replays_bit_exact_from_device_matrix(eng, oracle, n, d, mn, mx, threads):
    """Tensor-core initial matrix -> device loop; the oracle replays the SAME matrix."""
    x = synth.gaussian_mixture(n, d, mn, min(mx, 40), seed=n + d)
    eng.set_option("loop_threads", threads)
    try:
        with lw_only(eng):
            eng.load(x)
            eng.initial_distances(_lib.GRAM_TCGEN05_3XTF32, mx)
            m0 = eng.read_matrix()
            eng.nn_init()
            eng.merge_loop(mn, mx)
    finally:
        eng.set_option("loop_threads", 0)
    o = oracle.fast_cluster(x, mn, mx, flags=LW_EAGER, init_matrix=m0)
    _same_trace(eng.merge_trace(), o)
    cl = eng.build_clusters(mn)
    assert same_clusters(cl, o.clusters)
    sizes = [len(c) for c in cl]
    assert all(mn <= s <= mx for s in sizes)
    flat = np.concatenate(cl) if cl else np.zeros(0, np.int32)
    assert len(np.unique(flat)) == len(flat)  # every item at most once


# ---- the whole path, PerformClusteringWithConstraints (clustering.go:198-284) -------------

def test_config_a_matches_the_literal_reference_restatement(eng, oracle):
    """BASELINE config 1: N=1000 x 2048, 5/20.  Golden = literal restatement of the Go code."""
    g = load_golden("cfgA_1000x2048")
    ids = synth.item_ids(1000)
    cmap, ok = clustering.perform_clustering_with_constraints(g["x"], ids, 5, 20, engine=eng)
    assert ok
    want = golden_clusters(g)
    got = [np.array([int(s[4:]) for s in cmap[k]], np.int32) for k in range(len(cmap))]
    tr = eng.merge_trace()
    assert ari(got, want, 1000) == 1.0
    assert np.array_equal(tr.key_hi, g["key_hi"]) and np.array_equal(tr.key_lo, g["key_lo"])
    assert np.array_equal(tr.dist.view(np.uint32), g["dist"].view(np.uint32))  # the reference's own fp32 values
    assert same_clusters(got, want)
    _assert_reference_run(eng.stats())


def test_reference_error_behaviour(eng):
    x = np.zeros((3, 4), np.float32)
    assert clustering.perform_clustering_with_constraints(x, ["a", "b", "c"], 5, 20, engine=eng) == (None, False)
    x = np.zeros((7, 4), np.float32)
    assert clustering.perform_clustering_with_constraints(x, list("abcdefg"), 4, 5, engine=eng) == (None, False)
    rag = [[1.0, 2.0], [1.0], [0.0, 3.0]]
    assert clustering.perform_clustering_with_constraints(rag, list("abc"), 1, 2, engine=eng) == (None, False)
    x = np.random.default_rng(0).standard_normal((6, 3)).astype(np.float32)
    assert clustering.perform_clustering_with_constraints(x, list("abc"), 1, 2, engine=eng) == (None, False)


def test_trivial_sizes(eng):
    x = np.random.default_rng(1).standard_normal((5, 3)).astype(np.float32)
    cmap, ok = clustering.perform_clustering_with_constraints(x, list("abcde"), 1, 1, engine=eng)
    assert ok and cmap == {i: [c] for i, c in enumerate("abcde")}
    cmap, ok = clustering.perform_clustering_with_constraints(x[:1], ["only"], 1, 3, engine=eng)
    assert ok and cmap == {0: ["only"]}


def test_members_follow_the_reference_order(eng, oracle):
    # clustering.go:31,237: indices = clusters[i].Indices ++ clusters[j].Indices with i the larger position
    g = load_golden("n8_d2")
    eng.set_option("gram_mode", _lib.GRAM_EXACT_FP32)
    try:
        res = eng.cluster(g["x"], int(g["min_size"]), int(g["max_size"]))
    finally:
        eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)
    for a, b in zip(res.clusters, golden_clusters(g)):
        assert a.tolist() == b.tolist()


# ---- row-block sharding on ONE GPU: P virtual ranks in one cooperative launch -------------
# The same kernel path as the multi-GPU build -- merge_batch_kernel<multi> (the kernel bench.py --gpus N runs: exchange
# boxes, cross-rank barriers, peer row pointers, per-rank replicas of the slot table) with loop_mode 1, merge_loop_kernel
# with loop_mode 0 -- only the peers' memory happens to be on the same device.

@pytest.fixture
def knobs(eng):
    """Set merge-loop options for one test and restore the defaults afterwards."""
    def set_(**kw):
        for k, v in kw.items():
            eng.set_option(k, v)
    yield set_
    for k in ("virtual_ranks", "loop_blocks", "no_replica", "loop_mode", "exact", "mirror_init", "compact", "near_lists", "fast_start"):
        eng.set_option(k, 1 if k in ("virtual_ranks", "loop_mode", "exact", "mirror_init", "compact", "near_lists", "fast_start") else 0)


@pytest.mark.parametrize("name", SMALL_GOLDENS)
@pytest.mark.parametrize("ranks,loop_mode", [(2, 1), (3, 1), (8, 1), (2, 0), (8, 0)])
def test_virtual_shards_goldens_bit_exact(eng, oracle, knobs, name, ranks, loop_mode):
    g = load_golden(name)
    mn, mx = int(g["min_size"]), int(g["max_size"])
    o = oracle.fast_cluster(g["x"], mn, mx, flags=LW_EAGER, init_matrix=g["init_matrix"])
    knobs(virtual_ranks=ranks, loop_mode=loop_mode)
    eng.load(g["x"])
    eng.set_matrix(g["init_matrix"])
    eng.nn_init()
    eng.merge_loop(mn, mx)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(eng.build_clusters(mn), o.clusters)


@pytest.mark.parametrize("loop_mode", [1, 0])
@pytest.mark.parametrize("n,d,mn,mx,ranks,blocks,no_replica", [
    (3000, 64, 4, 12, 2, 0, 0), (3000, 64, 4, 12, 4, 8, 0), (3001, 64, 4, 12, 8, 3, 1), (5000, 32, 6, 8, 3, 16, 0),
    (2500, 100, 1, 2500, 2, 5, 1), (3000, 64, 4, 12, 1, 7, 1), (4000, 24, 2, 6, 1, 148, 0)])
def test_sharded_loop_replays_bit_exact(eng, oracle, knobs, n, d, mn, mx, ranks, blocks, no_replica, loop_mode):
    """Tensor-core initial matrix -> sharded device loop (virtual ranks, forced block counts; loop_mode 0: keys streamed
    from L2 instead of the shared-memory replica); the oracle replays the SAME matrix."""
    x = synth.gaussian_mixture(n, d, mn, min(mx, 40), seed=n + d)
    knobs(virtual_ranks=ranks, loop_blocks=blocks, no_replica=no_replica, loop_mode=loop_mode)
    with lw_only(eng):
        eng.load(x)
        eng.initial_distances(_lib.GRAM_TCGEN05_I8 if ranks % 2 else _lib.GRAM_TCGEN05_3XTF32, mx)
        m0 = eng.read_matrix()
        eng.nn_init()
        eng.merge_loop(mn, mx)
    o = oracle.fast_cluster(x, mn, mx, flags=LW_EAGER, init_matrix=m0)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(eng.build_clusters(mn), o.clusters)
    key, size = eng.read_slots()
    assert int((key >= 0).sum()) == o.n_final and int(size[key >= 0].sum()) == n


@pytest.mark.parametrize("loop_mode", [1, 0])
def test_sharded_duplicates_many_dry_rows(eng, oracle, knobs, loop_mode):
    """Dozens of rows lose all their cached partners at once (exact duplicates), on 4 virtual ranks: the batched sharded
    kernel with the horizon reproduces the reference arithmetic, the one-merge-per-iteration kernel (rescans served a few
    per iteration while lower bounds hold the merge back) the Lance-Williams oracle."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((400, 8)).astype(np.float32)
    x[50:120] = x[7]
    x[200:230] = x[9]
    o = oracle.fast_cluster(x, 1, 6, flags=0 if loop_mode == 1 else LW_EAGER)
    knobs(virtual_ranks=4, loop_blocks=2, loop_mode=loop_mode)
    eng.set_option("gram_mode", _lib.GRAM_EXACT_FP32)
    try:
        res = eng.cluster(x, 1, 6)
    finally:
        eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(res.clusters, o.clusters)


@pytest.mark.parametrize("loop_mode", [1, 0])
def test_sharded_staged_resume(eng, oracle, knobs, loop_mode):
    x = synth.gaussian_mixture(700, 48, 3, 10, seed=34)
    o = oracle.fast_cluster(x, 3, 10, flags=0 if loop_mode == 1 else LW_EAGER)
    knobs(virtual_ranks=2, loop_blocks=3, loop_mode=loop_mode)
    eng.load(x)
    eng.initial_distances(_lib.GRAM_EXACT_FP32, 10)
    eng.nn_init()
    for step in (1, 7, 100, 0, 13):
        eng.merge_loop(3, 10, step)
        hi, lo, d = eng.find_closest()
        t = eng.stats()["n_merges"]
        if t < o.n_merges:
            assert (hi, lo) == (int(o.key_hi[t]), int(o.key_lo[t])) and np.float32(d) == o.dist[t]
    eng.merge_loop(3, 10)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(eng.build_clusters(3), o.clusters)


@pytest.mark.parametrize("ranks,n,d,mn,mx,gram", [(2, 3000, 64, 4, 12, _lib.GRAM_TCGEN05_I8), (4, 2000, 2048, 10, 50, _lib.GRAM_TCGEN05_I8),
                                                  (8, 2500, 100, 1, 2500, _lib.GRAM_EXACT_FP32), (3, 5000, 32, 6, 8, _lib.GRAM_TCGEN05_3XTF32)])
def test_virtual_shards_equal_reference_arithmetic(eng, oracle, knobs, ranks, n, d, mn, mx, gram):
    """The sharded batched kernel (what bench.py --gpus N runs), P virtual ranks, whole path with the horizon: the merge
    trace is the reference-arithmetic one; every rank's queue, exact phase and exchange of the overflow / order flags ran."""
    x = synth.gaussian_mixture(n, d, mn, min(mx, 40), seed=ranks + n)
    o = oracle.fast_cluster(x, mn, mx, flags=0)
    knobs(virtual_ranks=ranks)
    eng.set_option("gram_mode", gram)
    try:
        res = eng.cluster(x, mn, mx)
    finally:
        eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(res.clusters, o.clusters)
    _assert_reference_run(res.stats)


def test_virtual_shards_more_candidates_than_the_exchange_box_holds(eng, oracle, knobs):
    """6 000 exact duplicate pairs on 2 virtual ranks: ~3 000 heads per rank sit below the stopper, more than a rank's
    region of the exchange box holds (kBatchXCand = 2 048).  Which candidates are dropped depends on the order of the
    atomics; the ranks then publish their minimum and merge the global minimum only -- the result must not depend on it."""
    rng = np.random.default_rng(11)
    base = (rng.standard_normal((6000, 8)) * 50).astype(np.float32)
    x = np.concatenate([base, base])[rng.permutation(12000)]
    o = oracle.fast_cluster(x, 1, 4, flags=0)
    knobs(virtual_ranks=2, gram_mode=_lib.GRAM_EXACT_FP32)
    try:
        res = eng.cluster(x, 1, 4)
    finally:
        eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(res.clusters, o.clusters)


# ---- BASELINE.json's full sizes --------------------------------------------------------------------

def _trace_digest(tr):
    import hashlib
    return hashlib.sha256(tr.key_hi.tobytes() + tr.key_lo.tobytes() + tr.dist.tobytes() + tr.size.tobytes()).hexdigest()


def test_n6000_equals_reference_arithmetic(eng, oracle):
    """N = 6,000 x 2048 (10/50), the size at which Lance-Williams values alone diverge from the reference (ARI 0.984):
    the product path must reproduce the reference-arithmetic oracle merge for merge -- unconditionally."""
    x = synth.gaussian_mixture(6000, 2048, 10, 50, seed=20241)
    o = oracle.fast_cluster(x, 10, 50, flags=0)
    res = eng.cluster(x, 10, 50)
    tr = eng.merge_trace()
    same = (tr.key_hi == o.key_hi) & (tr.key_lo == o.key_lo)
    assert same.all(), ("first divergence at merge", int(np.argmin(same)), "ARI", ari(res.clusters, o.clusters, 6000))
    _same_trace(tr, o)
    assert same_clusters(res.clusters, o.clusters)
    _assert_reference_run(res.stats)


def test_config_b_full_size_equals_reference_arithmetic(eng, oracle):
    """BASELINE config 2 (N=20,000 x 2048, 10/50) at full size against the oracle in REFERENCE arithmetic (all host
    cores): 18,800 merges, identical keys, distance bits and cluster lists; the tensor-core initial distances are
    within 1e-5 of the reference arithmetic on a sampled block."""
    n, d, mn, mx = synth.CONFIGS["B"]
    x = synth.gaussian_mixture(n, d, mn, mx, seed=20241)
    eng.load(x)
    eng.initial_distances(_lib.GRAM_TCGEN05_I8, mx)
    m0 = eng.read_matrix()
    rows = np.arange(0, n, 97)[:200]
    ref = oracle.initial_matrix(np.ascontiguousarray(x[rows]))
    sub = m0[np.ix_(rows, rows)]
    off = ~np.eye(len(rows), dtype=bool)
    assert float(np.max(np.abs(sub[off] - ref[off]) / ref[off])) <= RTOL
    del m0
    eng.nn_init()
    eng.merge_loop(mn, mx)
    tr = eng.merge_trace()
    o = oracle.fast_cluster(x, mn, mx, flags=0)
    same = (tr.key_hi == o.key_hi) & (tr.key_lo == o.key_lo)
    assert same.all(), ("first divergence at merge", int(np.argmin(same)))
    _same_trace(tr, o)
    cl = eng.build_clusters(mn)
    assert same_clusters(cl, o.clusters)
    assert all(mn <= len(c) <= mx for c in cl)
    _assert_reference_run(eng.stats())


def test_config_b_full_size_lance_williams_replay(eng, oracle):
    """The same size with "exact" = 0: the oracle's Lance-Williams mode replays the device's own tensor-core initial
    matrix, 18,800 merges bit for bit (loop-only parity of the arithmetic kept above the horizon)."""
    n, d, mn, mx = synth.CONFIGS["B"]
    x = synth.gaussian_mixture(n, d, mn, mx, seed=20241)
    with lw_only(eng):
        eng.load(x)
        eng.initial_distances(_lib.GRAM_TCGEN05_I8, mx)
        m0 = eng.read_matrix()
        eng.nn_init()
        eng.merge_loop(mn, mx)
        tr = eng.merge_trace()
    o = oracle.fast_cluster(x, mn, mx, flags=LW_EAGER, init_matrix=m0)
    _same_trace(tr, o)


def test_config_c_full_size_properties_and_sharded_identity(eng, knobs):
    """BASELINE config 3 (N=100,000 x 2048, 20/200) at full size: size-independent properties of the result and the
    horizon's guarantees (the oracle comparison at this size is scripts/replay_full.py, recorded under profiles/); with
    "exact" = 0 the row-block sharded loop (4 virtual ranks) must reproduce the single-rank Lance-Williams merge trace
    bit for bit (97,250 merges)."""
    n, d, mn, mx = synth.CONFIGS["C"]
    x = synth.gaussian_mixture(n, d, mn, mx, seed=20242)
    res = eng.cluster(x, mn, mx)
    tr = eng.merge_trace()
    st = res.stats
    _assert_reference_run(st)
    # sha256 of the merge trace (keys, distance bits, sizes) of the CPU oracle in REFERENCE arithmetic at this exact size:
    # oracle.fast_cluster(flags=0), 820 s on 16 host cores (scripts/replay_full.py C -> profiles/r02_replay_configC.txt)
    assert _trace_digest(tr) == "51f51453df136d2fe592f68de112c13745137c724d21f478d1281f8f7d9d7ce4"
    assert st["n_target"] == 2750 and st["n_merges"] == n - 2750 and not st["exhausted"]
    sizes = np.array([len(c) for c in res.clusters])
    assert sizes.min() >= mn and sizes.max() <= mx
    flat = np.concatenate(res.clusters)
    assert len(np.unique(flat)) == len(flat) and flat.min() >= 0 and flat.max() < n  # every item at most once
    # the trace is a valid dendrogram prefix: keys in range, each consumed once, sizes add up
    consumed = np.zeros(n + len(tr.key_hi), bool)
    size_of = np.concatenate([np.ones(n, np.int64), tr.size.astype(np.int64)])
    for t in range(len(tr.key_hi)):
        hi, lo = int(tr.key_hi[t]), int(tr.key_lo[t])
        assert lo < hi < n + t and not consumed[hi] and not consumed[lo]
        consumed[hi] = consumed[lo] = True
        assert size_of[n + t] == size_of[hi] + size_of[lo] <= mx
    # Ward heights never decrease by more than rounding (reducibility; exact in real arithmetic)
    dist = tr.dist.astype(np.float64)
    assert np.all(dist[1:] >= dist[:-1] * (1 - 1e-5))
    with lw_only(eng):
        eng.load(x)
        eng.run_resident(mn, mx)
        want = _trace_digest(eng.merge_trace())
        knobs(virtual_ranks=4)
        eng.load(x)
        eng.run_resident(mn, mx)
        assert _trace_digest(eng.merge_trace()) == want


@pytest.mark.parametrize("mn,want,exhausted", [
    (2, "12a460c80c3263b8a5acf911620e40e897f3cea84349fdc877bc9d94bd3cb334", False),
    (6, "24327edbc359afde85a919e226988cc5228c3a65c7b52b12859aae6fd98a866e", True)])
def test_config_e_full_size_equals_the_recorded_oracle_digest(eng, mn, want, exhausted):
    """BASELINE config 5 (N=50,000 x 2148: 2048-d image block + 100-d label one-hot, maxSize=8) at full size, minSize 2 and 6
    (the latter ends by exhaustion): the sha256 of the merge trace equals the one of the CPU oracle in reference arithmetic
    at this exact size (171 s / 222 s on 16 host cores: scripts/replay_full.py E:2 E:6:1 -> profiles/r02_replay_configE.txt)."""
    n, d, _, mx = synth.CONFIGS["E"]
    x = synth.combined_features(n, 2048, d - 2048, 2, 8, seed=20244)
    res = eng.cluster(x, mn, mx)
    assert _trace_digest(eng.merge_trace()) == want
    assert bool(res.stats["exhausted"]) == exhausted
    _assert_reference_run(res.stats)
    sizes = np.array([len(c) for c in res.clusters])
    assert sizes.min() >= mn and sizes.max() <= mx


# ---- K1, int8 path: exact integer tensor-core Gram (gram_i8.cu) -------------------------------------------------

def _gram_case_mode(eng, oracle, x, mode):
    eng.load(x)
    eng.initial_distances(mode)
    m = eng.read_matrix()
    ref = oracle.initial_matrix(x)
    assert np.array_equal(m, m.T)
    assert np.all(np.diag(m) == 0)
    off = ~np.eye(len(x), dtype=bool)
    return float((np.abs(m[off] - ref[off]) / ref[off]).max())


@pytest.mark.parametrize("n,d,relu", [(256, 64, False), (700, 2048, False), (700, 2048, True), (333, 2148, False),
                                      (1000, 96, False), (129, 40, False), (130, 300, True)])
def test_i8_gram_within_tolerance(eng, oracle, n, d, relu):
    x = synth.gaussian_mixture(n, d, 5, 20, seed=100 + n + d, relu_like=relu)
    assert _gram_case_mode(eng, oracle, x, _lib.GRAM_TCGEN05_I8) <= RTOL


def test_i8_gram_config_e_like(eng, oracle):
    x = synth.combined_features(600, 2048, 100, 2, 8, seed=7)
    assert _gram_case_mode(eng, oracle, x, _lib.GRAM_TCGEN05_I8) <= RTOL


def test_i8_gram_full_path_replays_bit_exact(eng, oracle):
    n, d, mn, mx = 3000, 64, 4, 12
    x = synth.gaussian_mixture(n, d, mn, mx, seed=n + d)
    with lw_only(eng):
        eng.load(x)
        eng.initial_distances(_lib.GRAM_TCGEN05_I8, mx)
        m0 = eng.read_matrix()
        eng.nn_init()
        eng.merge_loop(mn, mx)
    o = oracle.fast_cluster(x, mn, mx, flags=LW_EAGER, init_matrix=m0)
    _same_trace(eng.merge_trace(), o)


# ---- K3b: batched merge loop (merge_batch.cu, the default on one GPU) vs the one-merge-per-iteration loop ----------

@pytest.mark.parametrize("name", SMALL_GOLDENS)
def test_sequential_loop_mode_goldens_bit_exact(eng, oracle, knobs, name):
    """loop_mode=0 keeps the one-merge-per-iteration kernel covered on a single rank (the default is batched)."""
    g = load_golden(name)
    mn, mx = int(g["min_size"]), int(g["max_size"])
    o = oracle.fast_cluster(g["x"], mn, mx, flags=LW_EAGER, init_matrix=g["init_matrix"])
    knobs(loop_mode=0)
    eng.load(g["x"])
    eng.set_matrix(g["init_matrix"])
    eng.nn_init()
    eng.merge_loop(mn, mx)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(eng.build_clusters(mn), o.clusters)


@pytest.mark.parametrize("n,d,mn,mx", [(3000, 64, 4, 12), (6000, 256, 10, 50), (5000, 32, 6, 8)])
def test_batched_and_sequential_loops_agree(eng, knobs, n, d, mn, mx):
    """The batch rule only ever takes merges the sequential algorithm would take next, in the same order."""
    x = synth.gaussian_mixture(n, d, mn, mx, seed=7 * n + d)
    digests, iters = [], []
    with lw_only(eng):  # the one-merge-per-iteration loop keeps Lance-Williams values only
        for mode in (1, 0):
            knobs(loop_mode=mode)
            eng.load(x)
            res = eng.run_resident(mn, mx)
            digests.append(_trace_digest(eng.merge_trace()))
            iters.append(res.stats["n_merges"])
    assert digests[0] == digests[1]


@pytest.mark.parametrize("ratio,mirror", [(0.5, 1), (0.7, 0), (0.9, 1)])
def test_compaction_does_not_change_the_trace(eng, knobs, ratio, mirror):
    """K4: renumbering the live clusters and moving the matrix (any ratio, with or without the initial mirror pass) leaves
    the merge trace bit-identical -- in Lance-Williams arithmetic (every stored value matters) and with the horizon."""
    x = synth.gaussian_mixture(9000, 96, 4, 12, seed=19)
    for exact in (0, 1):
        digests, ncomp = [], []
        # no compaction; the one-pass tile kernel; compact_rows + mirror_lower (what a sharded run uses)
        for compact, tiles in ((0, 1), (1, 1), (1, 0)):
            knobs(exact=exact, compact=compact, mirror_init=mirror)
            eng.set_option("compact_ratio", ratio)
            eng.set_option("compact_tiles", tiles)
            try:
                res = eng.cluster(x, 4, 12)
            finally:
                eng.set_option("compact_ratio", 0.7)
                eng.set_option("compact_tiles", 1)
            digests.append(_trace_digest(eng.merge_trace()))
            ncomp.append(res.stats["n_compactions"])
        assert digests[0] == digests[1] == digests[2]
        assert ncomp[0] == 0 and ncomp[1] >= 1 and ncomp[2] == ncomp[1]


@pytest.mark.parametrize("n,d,mn,mx,ranks", [(9000, 96, 4, 12, 1), (6000, 2048, 10, 50, 1), (5000, 32, 6, 8, 1), (6000, 64, 4, 12, 3)])
def test_near_lists_do_not_change_the_trace(eng, knobs, n, d, mn, mx, ranks):
    """Partner lists re-selected from the rows' near lists (every pair at or below the horizon, near.cu) instead of full row
    scans: same trace, far fewer bytes; also across horizon raises, compactions, exhaustion and on virtual ranks."""
    x = synth.gaussian_mixture(n, d, mn, mx, seed=23 + n)
    digests = []
    for near in (0, 1):
        knobs(near_lists=near, virtual_ranks=ranks)
        res = eng.cluster(x, mn, mx)
        digests.append(_trace_digest(eng.merge_trace()))
        _assert_reference_run(res.stats)
    assert digests[0] == digests[1]


@pytest.mark.parametrize("n,d,mn,mx", [(6000, 2048, 10, 50), (5000, 32, 6, 8), (3000, 64, 1, 4), (9000, 96, 4, 12)])
def test_start_without_the_first_sweep_does_not_change_the_trace(eng, knobs, n, d, mn, mx):
    """With the horizon and near lists on, K2's only product that outlives the first horizon is the global minimum, which the
    mirror pass after K1 collects: the run that skips K2 (default) and the one that does not give the same trace."""
    x = synth.gaussian_mixture(n, d, mn, mx, seed=41 + n)
    digests, launches = [], []
    for fast in (0, 1):
        knobs(fast_start=fast)
        res = eng.cluster(x, mn, mx)
        digests.append(_trace_digest(eng.merge_trace()))
        launches.append(res.stats["kernel_launches"])
        _assert_reference_run(res.stats)
    assert digests[0] == digests[1]
    assert launches[1] < launches[0]  # (no first sweep, no first loop launch that only reports the minimum)


def test_barrier_timeout_is_reported_and_the_context_survives(eng, oracle):
    """A block that never arrives at a grid barrier (test hook loop_debug = 77) must not hang the GPU or poison the CUDA
    context: the other blocks give the barrier up after 2^24 polls, the call fails with the library's internal-error code,
    and the same engine clusters correctly afterwards."""
    x = synth.gaussian_mixture(3000, 64, 4, 12, seed=5)
    o = oracle.fast_cluster(x, 4, 12, flags=0)
    eng.set_option("loop_debug", 77)
    try:
        with pytest.raises(clustering.EngineError) as ei:
            eng.cluster(x, 4, 12)
        assert ei.value.code == _lib.IC_ERR_INTERNAL
    finally:
        eng.set_option("loop_debug", 0)
    res = eng.cluster(x, 4, 12)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(res.clusters, o.clusters)


def test_batched_more_disjoint_pairs_than_the_batch_capacity(eng, oracle, knobs):
    """1 500 exact duplicate pairs: every pair is a head at distance 0 and none conflicts, so far more than
    kMaxBatch (512) pairs sit below the stopper; the batch is cut to a prefix by bisection on the packed value."""
    rng = np.random.default_rng(11)
    base = (rng.standard_normal((1500, 8)) * 50).astype(np.float32)
    x = np.concatenate([base, base])[rng.permutation(3000)]
    o = oracle.fast_cluster(x, 1, 4, flags=0)
    knobs(gram_mode=_lib.GRAM_EXACT_FP32)
    try:
        res = eng.cluster(x, 1, 4)
    finally:
        eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)
    _same_trace(eng.merge_trace(), o)
    assert same_clusters(res.clusters, o.clusters)


def test_batched_loop_exhaustion_and_tight_max(eng, oracle, knobs):
    """maxSize = 8 with minSize = 6 ends by exhaustion (clustering.go:222-225): the batched loop must stop at the
    same merge as the oracle."""
    x = synth.gaussian_mixture(2000, 32, 6, 8, seed=5)
    o = oracle.fast_cluster(x, 6, 8, flags=0)
    knobs(gram_mode=_lib.GRAM_EXACT_FP32)
    try:
        res = eng.cluster(x, 6, 8)
    finally:
        eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)
    _same_trace(eng.merge_trace(), o)
    assert bool(res.stats["exhausted"]) == o.exhausted
    assert same_clusters(res.clusters, o.clusters)


def test_batched_loop_stats_and_mode_guard(eng, knobs):
    """ic_stats reports the loop that ran and its iterations; the loop mode cannot change while the matrix holds the lower
    triangle only (K1 of the batched path without the mirror pass: the one-merge-per-iteration loop needs both)."""
    x = synth.gaussian_mixture(1500, 64, 4, 12, seed=77)
    knobs(exact=0, mirror_init=0)  # the two loops share Lance-Williams arithmetic only
    res = eng.cluster(x, 4, 12)
    assert res.stats["loop_mode"] == 1
    assert 0 < res.stats["n_iterations"] < res.stats["n_merges"]
    eng.load(x)
    eng.initial_distances(_lib.GRAM_TCGEN05_I8, 12)
    m0 = eng.read_matrix()
    assert np.array_equal(m0, m0.T)  # K1 stored the lower triangle only; the caller still gets the symmetric matrix
    eng.nn_init()
    eng.set_option("loop_mode", 0)
    try:
        with pytest.raises(Exception):
            eng.merge_loop(4, 12)
    finally:
        eng.set_option("loop_mode", 1)
    knobs(loop_mode=0)
    eng.load(x)
    res0 = eng.run_resident(4, 12)
    assert res0.stats["loop_mode"] == 0 and res0.stats["n_iterations"] >= res0.stats["n_merges"]
    assert same_clusters(res0.clusters, res.clusters)
