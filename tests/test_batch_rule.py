"""CPU check of the claim the batched device loop rests on: taking, per iteration, every row head
below the first pair that touches an earlier pair's cluster reproduces the SEQUENTIAL merge sequence bit for bit.
The numpy restatement of the rule (oracle/batch_rule.py) runs against the oracle's Lance-Williams mode, which is pinned
against the literal restatement of clustering.go by tests/test_oracle_*.py."""
import numpy as np
import pytest

from imageclust_b200 import clustering, synth
from oracle import batch_rule
from tests.helpers import golden_names, load_golden

LW_EAGER = 3  # oracle.FAST_EAGER | oracle.FAST_LW
SMALL = [n for n in golden_names() if n != "cfgA_1000x2048"]


@pytest.mark.parametrize("name", SMALL)
def test_batch_rule_reproduces_the_sequential_trace_on_goldens(oracle, name):
    g = load_golden(name)
    mn, mx = int(g["min_size"]), int(g["max_size"])
    o = oracle.fast_cluster(g["x"], mn, mx, flags=LW_EAGER, init_matrix=g["init_matrix"])
    hi, lo, d, s, batches = batch_rule.batched_cluster(g["init_matrix"], int(g["n_target"]), mx)
    assert len(hi) == o.n_merges
    assert np.array_equal(hi, o.key_hi) and np.array_equal(lo, o.key_lo)
    assert np.array_equal(d.view(np.uint32), o.dist.view(np.uint32)) and np.array_equal(s, o.size)


@pytest.mark.parametrize("n,d,mn,mx,dup", [(300, 16, 2, 5, 0), (260, 8, 6, 8, 0), (240, 4, 1, 7, 60), (200, 3, 1, 200, 0)])
def test_batch_rule_random_inputs_with_ties_rejections_and_exhaustion(oracle, n, d, mn, mx, dup):
    rng = np.random.default_rng(n + d)
    x = rng.standard_normal((n, d)).astype(np.float32)
    if dup:
        x[rng.integers(0, n, dup)] = x[rng.integers(0, n, dup)]  # exact ties
    o = oracle.fast_cluster(x, mn, mx, flags=LW_EAGER)
    m0 = oracle.initial_matrix(x)
    n_target = clustering.calculate_optimal_clusters(n, mn, mx)[0]
    hi, lo, dist, s, batches = batch_rule.batched_cluster(m0, n_target, mx)
    assert len(hi) == o.n_merges
    assert np.array_equal(hi, o.key_hi) and np.array_equal(lo, o.key_lo)
    assert np.array_equal(dist.view(np.uint32), o.dist.view(np.uint32)) and np.array_equal(s, o.size)
    assert batches.sum() == o.n_merges and batches.max() > 1  # the rule really batches


def test_batch_rule_capacity_cut_is_a_prefix(oracle):
    """Cutting a batch to a prefix (the kernel's 512-merge capacity) must not change the sequence."""
    x = synth.gaussian_mixture(220, 8, 2, 6, seed=3)
    o = oracle.fast_cluster(x, 2, 6, flags=LW_EAGER)
    m0 = oracle.initial_matrix(x)
    n_target = clustering.calculate_optimal_clusters(220, 2, 6)[0]
    for cap in (1, 2, 3):
        hi, lo, dist, s, batches = batch_rule.batched_cluster(m0, n_target, 6, max_batch=cap)
        assert np.array_equal(hi, o.key_hi) and np.array_equal(lo, o.key_lo)
        assert np.array_equal(dist.view(np.uint32), o.dist.view(np.uint32))
        assert batches.max() <= cap


# ---- property test: integer-grid points (masses of exact ties), random constraints ------------------------------
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(4, 28), d=st.integers(1, 3), grid=st.integers(2, 5), mn=st.integers(1, 4), extra=st.integers(0, 6),
       seed=st.integers(0, 10**6))
def test_batch_rule_property_exact_ties(oracle, n, d, grid, mn, extra, seed):
    """Points on a small integer grid: many pairs at exactly the same distance, duplicates, chains of equal heads.  The
    tie-break (d, key_hi, key_lo) and the 'ties go to the older pair' argument of the batch rule are what is tested."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, grid, size=(n, d)).astype(np.float32)
    mx = mn + extra
    n_target, err = clustering.calculate_optimal_clusters(n, mn, mx)
    if err is not None:
        return
    o = oracle.fast_cluster(x, mn, mx, flags=LW_EAGER)
    m0 = oracle.initial_matrix(x)
    hi, lo, dist, s, batches = batch_rule.batched_cluster(m0, n_target, mx)
    assert len(hi) == o.n_merges
    assert np.array_equal(hi, o.key_hi) and np.array_equal(lo, o.key_lo)
    assert np.array_equal(dist.view(np.uint32), o.dist.view(np.uint32)) and np.array_equal(s, o.size)
