"""CPU proof of the DEVICE's merge-loop algorithm against the reference arithmetic.

``oracle.device_cluster`` (oracle/ward_device.c) restates what merge_batch.cu + refine.cu do -- Lance-Williams values above
a horizon, WardDistance of the fp32 centroids (clustering.go:83-86,136-157) at or below it, batches of consecutive merges --
and must reproduce ``oracle.fast_cluster(flags=0)``, which is bit-identical to the literal restatement of clustering.go
(tests/test_oracle_fast.py), merge for merge and bit for bit.  No GPU involved."""
import numpy as np
import pytest

from imageclust_b200 import synth


def _same(a, b):
    return (a.n_merges == b.n_merges and np.array_equal(a.key_hi, b.key_hi) and np.array_equal(a.key_lo, b.key_lo)
            and np.array_equal(a.dist.view(np.uint32), b.dist.view(np.uint32)) and np.array_equal(a.size, b.size)
            and len(a.clusters) == len(b.clusters) and all(np.array_equal(p, q) for p, q in zip(a.clusters, b.clusters)))


@pytest.mark.parametrize("n,d,mn,mx,seed", [(1500, 2048, 10, 50, 20241), (2000, 64, 4, 12, 3), (1200, 32, 6, 8, 5),
                                            (900, 100, 1, 900, 7)])
@pytest.mark.parametrize("delta_cut", [0.0, 8e-6])
def test_device_algorithm_reproduces_the_reference_arithmetic(oracle, n, d, mn, mx, seed, delta_cut):
    x = synth.gaussian_mixture(n, d, mn, min(mx, 40), seed=seed)
    ref = oracle.fast_cluster(x, mn, mx, flags=0)
    got, ds = oracle.device_cluster(x, mn, mx, delta_cut=delta_cut)
    assert _same(got, ref), ds
    assert ds["n_violations"] == 0 and ds["max_filter_err"] < 3e-5
    assert ds["n_iterations"] < max(ref.n_merges, 1) or ref.n_merges <= 1  # it does batch
    assert ref.exhausted == got.exhausted


def test_horizon_repairs_an_approximate_initial_matrix(oracle):
    """The tensor-core Gram is off by a few 1e-6: the sweep at the first horizon replaces every value that can matter."""
    x = synth.gaussian_mixture(1500, 2048, 10, 50, seed=20241)
    ref = oracle.fast_cluster(x, 10, 50, flags=0)
    m0 = oracle.initial_matrix(x)
    rng = np.random.default_rng(1)
    mp = np.tril((m0.astype(np.float64) * (1 + rng.uniform(-8e-6, 8e-6, size=m0.shape))).astype(np.float32), -1)
    got, ds = oracle.device_cluster(x, 10, 50, init_matrix=mp + mp.T)
    assert _same(got, ref), ds
    assert ds["n_exact"] > 0 and ds["max_filter_err"] < 3e-5


def test_lance_williams_alone_does_not(oracle):
    """Why the horizon exists: without it (oracle LW mode == the device with option "exact" = 0) near-ties flip."""
    x = synth.gaussian_mixture(3000, 2048, 10, 50, seed=20241)
    ref = oracle.fast_cluster(x, 10, 50, flags=0)
    lw = oracle.fast_cluster(x, 10, 50, flags=3)
    assert not (np.array_equal(ref.key_hi, lw.key_hi) and np.array_equal(ref.key_lo, lw.key_lo))
    got, _ = oracle.device_cluster(x, 10, 50)
    assert _same(got, ref)


def test_duplicates_and_small_horizon_factor(oracle):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((400, 8)).astype(np.float32)
    x[50:120] = x[7]
    x[200:230] = x[9]
    ref = oracle.fast_cluster(x, 1, 6, flags=0)
    for f in (1.25, 1.02):
        got, ds = oracle.device_cluster(x, 1, 6, horizon_factor=f)
        assert _same(got, ref), (f, ds)
