"""oracle_fast (NN cache, slot reuse, OpenMP) must be bit-identical to the literal
oracle; its EAGER / LW modes validate the device algorithm's reasoning on the CPU."""
import numpy as np
import pytest

from imageclust_b200 import synth
from tests.helpers import ari, golden_clusters, golden_names, load_golden, same_clusters

F32 = np.float32


def _same_trace(a, b):
    return (np.array_equal(a.key_hi, b.key_hi) and np.array_equal(a.key_lo, b.key_lo)
            and np.array_equal(a.dist, b.dist) and np.array_equal(a.size, b.size))


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("flags", [0, 1])
def test_fast_matches_golden(oracle, name, flags):
    g = load_golden(name)
    r = oracle.fast_cluster(g["x"], int(g["min_size"]), int(g["max_size"]), flags=flags)
    assert r.ok
    assert np.array_equal(r.key_hi, g["key_hi"]) and np.array_equal(r.key_lo, g["key_lo"])
    assert np.array_equal(r.dist, g["dist"]) and np.array_equal(r.size, g["size"])
    assert same_clusters(r.clusters, golden_clusters(g))
    assert int(r.exhausted) == int(g["exhausted"]) and r.n_final == int(g["n_final"])
    if flags == 0:   # lazy mode reproduces even the rejection count (clustering.go:228-234)
        assert r.n_rejections == int(g["n_rejections"])
    else:
        assert r.n_rejections == 0


@pytest.mark.parametrize("seed,n,d,mn,mx,dup", [
    (0, 300, 16, 2, 5, 0), (1, 400, 8, 6, 8, 0), (2, 350, 4, 1, 7, 40), (3, 500, 32, 3, 6, 0),
    (4, 257, 3, 2, 2, 0), (5, 300, 6, 1, 300, 0),
])
def test_fast_equals_literal(oracle, seed, n, d, mn, mx, dup):
    rng = np.random.default_rng(100 + seed)
    x = rng.standard_normal((n, d)).astype(F32)
    if dup:
        x[rng.integers(0, n, dup)] = x[rng.integers(0, n, dup)]
    lit = oracle.literal_cluster(x, mn, mx)
    for flags in (0, oracle.FAST_EAGER):
        for threads in (1, 3):
            f = oracle.fast_cluster(x, mn, mx, flags=flags, n_threads=threads)
            assert f.ok == lit.ok
            if not lit.ok:
                continue
            assert _same_trace(f, lit)
            assert same_clusters(f.clusters, lit.clusters)
            assert f.exhausted == lit.exhausted and f.n_final == lit.n_final
            if flags == 0:
                assert f.n_rejections == lit.n_rejections


def test_fast_equals_literal_config_like(oracle):
    # a 2048-wide case at a size the literal oracle finishes in ~1 s
    x = synth.gaussian_mixture(400, 2048, 5, 20, seed=3)
    lit = oracle.literal_cluster(x, 5, 20)
    f = oracle.fast_cluster(x, 5, 20)
    assert _same_trace(f, lit) and same_clusters(f.clusters, lit.clusters)


@pytest.mark.parametrize("lw_flags", [3, 7])
def test_lance_williams_mode_tracks_centroid_mode(oracle, lw_flags):
    """The device replaces the centroid recompute (clustering.go:83-86) by the
    Lance-Williams recurrence.  On tie-free data the merge sequence must be the
    same and merge distances within 1e-5 relative (BASELINE north_star)."""
    x = synth.gaussian_mixture(600, 64, 4, 12, seed=21)
    ref = oracle.fast_cluster(x, 4, 12)
    lw = oracle.fast_cluster(x, 4, 12, flags=lw_flags)
    assert ref.n_merges == lw.n_merges
    assert np.array_equal(ref.key_hi, lw.key_hi) and np.array_equal(ref.key_lo, lw.key_lo)
    np.testing.assert_allclose(lw.dist, ref.dist, rtol=1e-5)
    assert same_clusters(ref.clusters, lw.clusters)
    assert ari(ref.clusters, lw.clusters, 600) == 1.0


def test_init_matrix_replay(oracle):
    x = synth.gaussian_mixture(200, 32, 3, 9, seed=5)
    m = oracle.initial_matrix(x)
    lit = oracle.literal_cluster(x, 3, 9, want_matrices=True)
    assert np.array_equal(m, lit.init_matrix)
    f = oracle.fast_cluster(x, 3, 9, init_matrix=m)
    assert _same_trace(f, lit)
