"""Pins the CPU oracle (oracle/ward_literal.c) -- the reference has no tests of its
own, so the pins are: hand-computed cases, an independent numpy restatement,
scipy's Ward linkage, and the committed golden vectors."""
import numpy as np
import pytest

from imageclust_b200 import synth
from oracle import numpy_literal as NL
from tests.helpers import golden_clusters, golden_names, load_golden, same_clusters

F32 = np.float32


def test_calculate_optimal_clusters_branches(oracle):
    # clustering.go:168-186
    assert oracle.optimal_clusters(1000, 5, 20) == (125, 0)          # (50+200)/2
    assert oracle.optimal_clusters(20000, 10, 50) == (1200, 0)
    assert oracle.optimal_clusters(100000, 20, 200) == (2750, 0)
    assert oracle.optimal_clusters(50000, 2, 8) == (15625, 0)
    assert oracle.optimal_clusters(50000, 6, 8) == (7291, 0)
    assert oracle.optimal_clusters(10, 5, 5) == (2, 0)               # lo == hi
    assert oracle.optimal_clusters(3, 5, 20)[1] == oracle.ERR_TOO_FEW  # :169
    assert oracle.optimal_clusters(7, 4, 5)[1] == oracle.ERR_UNSAT     # lo=2 > hi=1, :175
    assert oracle.optimal_clusters(0, 0, 5)[1] == oracle.ERR_BAD_ARG
    assert oracle.optimal_clusters(5, 1, 0)[1] == oracle.ERR_BAD_ARG
    for n, mn, mx in [(17, 1, 1), (17, 3, 6), (99, 7, 9), (6, 6, 6), (1, 1, 1)]:
        got = oracle.optimal_clusters(n, mn, mx)
        want = NL.calculate_optimal_clusters(n, mn, mx)
        assert (got[1] == 0) == (want[1] is None)
        if want[1] is None:
            assert got[0] == want[0]


def test_dot_is_sequential_fp32(oracle):
    # clustering.go:152-155: order matters in fp32
    a = np.array([1e8, 1.0, -1e8, 1.0], F32)
    ones = np.ones(4, F32)
    assert oracle.dot_f32(a, ones) == 1.0           # ((1e8+1)-1e8)+1 with fp32 rounding
    rng = np.random.default_rng(1)
    for d in (1, 7, 64, 2048):
        u = rng.standard_normal(d).astype(F32)
        v = rng.standard_normal(d).astype(F32)
        assert oracle.dot_f32(u, v) == float(NL.dot_float32(u, v))


def test_ward_distance_and_centroid(oracle):
    # clustering.go:136-145 and :39
    rng = np.random.default_rng(2)
    for d in (2, 33, 2048):
        a = rng.standard_normal(d).astype(F32)
        b = rng.standard_normal(d).astype(F32)
        for sa, sb in ((1, 1), (3, 2), (17, 40)):
            assert oracle.ward_distance(a, sa, b, sb) == float(NL.ward_distance(a, sa, b, sb))
            assert np.array_equal(oracle.merge_centroid(a, sa, b, sb), NL.merge_centroid(a, sa, b, sb))
    # singleton weight is exactly 1/2
    assert oracle.ward_distance(np.array([3.0], F32), 1, np.array([0.0], F32), 1) == 4.5


def test_line6_by_hand(oracle):
    # points 0,1,3,6,10,15 on a line; min=1,max=3 -> lo=2, hi=6 -> target 4 -> two merges
    x = np.array([[0], [1], [3], [6], [10], [15]], F32)
    r = oracle.literal_cluster(x, 1, 3)
    assert r.ok and r.n_target == 4 and r.n_merges == 2
    # merge 1: positions (1,0), d = 0.5 * 1^2
    assert (r.pos_i[0], r.pos_j[0], r.dist[0], r.size[0]) == (1, 0, 0.5, 2)
    # slice is now [3,6,10,15,{1,0}]; {1,0} has centroid 0.5, size 2; d to x=3: (2*1/3)*2.5^2
    assert (r.pos_i[1], r.pos_j[1], r.size[1]) == (4, 0, 3)
    assert r.dist[1] == F32(F32(2) / F32(3)) * F32(6.25)
    assert [list(c) for c in r.clusters] == [[3], [4], [5], [1, 0, 2]]   # members hi ++ lo, :31
    assert (r.key_hi[1], r.key_lo[1]) == (6, 2)


def test_rejection_marks_and_min_size_drop(oracle):
    # two tight pairs far apart plus a straggler; max=2 forbids growing a pair
    x = np.array([[0, 0], [0.1, 0], [5, 0], [5.1, 0], [2.4, 0]], F32)
    r = oracle.literal_cluster(x, 2, 2)   # lo=3, hi=2 -> unsat
    assert not r.ok and r.rc == oracle.ERR_UNSAT
    r = oracle.literal_cluster(x, 1, 2)   # lo=3, hi=5 -> 4: one merge
    assert r.ok and r.n_merges == 1
    r = oracle.literal_cluster(np.vstack([x, [[9, 9]]]).astype(F32), 2, 3)  # N=6: lo=2, hi=3 -> 2
    assert r.ok and r.n_rejections > 0
    got = sorted(sorted(c.tolist()) for c in r.clusters)
    assert all(2 <= len(c) <= 3 for c in got)
    seen = [i for c in got for i in c]
    assert len(seen) == len(set(seen))


def test_exhaustion_exit(oracle):
    # clustering.go:222-225: min ~ max makes the loop run out of admissible pairs
    x = synth.gaussian_mixture(64, 16, 6, 8, seed=12)
    r = oracle.literal_cluster(x, 6, 8)
    assert r.ok and r.exhausted and r.n_final > r.n_target
    assert all(6 <= len(c) <= 8 for c in r.clusters)


@pytest.mark.parametrize("seed,n,d,mn,mx,dup", [
    (0, 40, 8, 2, 5, 0), (1, 60, 16, 3, 6, 0), (2, 50, 4, 1, 4, 10), (3, 64, 16, 6, 8, 0),
    (4, 30, 3, 1, 30, 0), (5, 45, 5, 5, 5, 5), (6, 33, 2, 2, 3, 0),
])
def test_c_literal_equals_numpy_literal(oracle, seed, n, d, mn, mx, dup):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(F32)
    if dup:
        x[rng.integers(0, n, dup)] = x[rng.integers(0, n, dup)]   # exact ties
    r = oracle.literal_cluster(x, mn, mx, want_matrices=True)
    q = NL.perform_clustering_with_constraints(x, mn, mx)
    assert r.ok == (q is not None)
    if q is None:
        return
    assert r.n_merges == len(q["trace"]) and r.n_rejections == q["rejections"]
    assert r.exhausted == q["exhausted"]
    tr = q["trace"]
    assert r.key_hi.tolist() == [t[0] for t in tr] and r.key_lo.tolist() == [t[1] for t in tr]
    assert r.pos_i.tolist() == [t[2] for t in tr] and r.pos_j.tolist() == [t[3] for t in tr]
    assert r.dist.tolist() == [t[4] for t in tr]
    assert np.array_equal(r.init_matrix, q["init_matrix"])
    assert np.array_equal(r.final_matrix, q["final_matrix"])
    assert r.final_keys.tolist() == q["final_keys"]
    assert same_clusters(r.clusters, q["clusters"])


def test_against_scipy_ward(oracle):
    """Unconstrained, tie-free: scipy height h satisfies h^2 / 2 == d_ref and the
    merge order is the same."""
    from scipy.cluster.hierarchy import linkage
    rng = np.random.default_rng(5)
    n = 80
    x = rng.standard_normal((n, 6)).astype(F32)
    r = oracle.literal_cluster(x, 1, n)        # lo=1, hi=n -> target (1+n)//2
    z = linkage(x.astype(np.float64), "ward")
    m = r.n_merges
    assert m == n - (1 + n) // 2
    np.testing.assert_allclose(z[:m, 2] ** 2 / 2, r.dist, rtol=2e-5)
    # same pairs: scipy ids are item index / n + merge index == our keys
    got = [tuple(sorted(p)) for p in zip(r.key_hi.tolist(), r.key_lo.tolist())]
    want = [tuple(sorted((int(a), int(b)))) for a, b in z[:m, :2]]
    assert got == want


@pytest.mark.parametrize("name", golden_names())
def test_golden_vectors(oracle, name):
    g = load_golden(name)
    r = oracle.literal_cluster(g["x"], int(g["min_size"]), int(g["max_size"]), want_matrices="init_matrix" in g)
    assert r.ok
    assert np.array_equal(r.key_hi, g["key_hi"]) and np.array_equal(r.key_lo, g["key_lo"])
    assert np.array_equal(r.dist, g["dist"]) and np.array_equal(r.size, g["size"])
    assert r.n_rejections == int(g["n_rejections"]) and int(r.exhausted) == int(g["exhausted"])
    assert same_clusters(r.clusters, golden_clusters(g))
    if "init_matrix" in g:
        assert np.array_equal(r.init_matrix, g["init_matrix"])
        assert np.array_equal(r.final_matrix, g["final_matrix"])
