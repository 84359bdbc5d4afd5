"""Dendrogram export and the caller-side report (8f-2).  The reference has no on-disk format; the export
uses scipy's linkage layout: keys are item indices or n + t for the cluster of merge t (scipy's own numbering), heights
are sqrt(2 d) because Ward's d = |a||b|/(|a|+|b|) * ||c_a - c_b||^2 is half of scipy's squared Ward distance."""
import numpy as np
import pytest
from scipy.cluster.hierarchy import linkage

from imageclust_b200 import clustering

RTOL = 1e-5


def _unconstrained(n):
    # minSize = 1, maxSize = n: n_target = (1 + n) // 2, no pair is ever inadmissible -> a prefix of plain Ward linkage
    return 1, n


def test_trace_to_linkage_matches_scipy_ward_prefix(oracle):
    rng = np.random.default_rng(8)
    n = 120
    x = rng.standard_normal((n, 5)).astype(np.float32)
    mn, mx = _unconstrained(n)
    o = oracle.literal_cluster(x, mn, mx)
    z = clustering.trace_to_linkage(o)
    want = linkage(x.astype(np.float64), "ward")[: len(z)]
    assert len(z) == n - (1 + n) // 2
    np.testing.assert_allclose(z[:, 2], want[:, 2], rtol=2e-5)
    assert np.array_equal(z[:, 3], want[:, 3])
    # scipy lists the smaller id first as well; tie-free data: the same pairs in the same order
    assert np.array_equal(np.sort(z[:, :2], axis=1), np.sort(want[:, :2], axis=1))
    assert np.all(z[:, 0] < z[:, 1])


def test_clustering_report_lists_dropped_items(oracle):
    x = np.random.default_rng(3).standard_normal((90, 4)).astype(np.float32)
    o = oracle.literal_cluster(x, 4, 6)
    rep = clustering.clustering_report(90, o.clusters)
    kept = np.concatenate(o.clusters) if o.clusters else np.zeros(0, np.int32)
    assert len(rep["dropped_items"]) + len(kept) == 90
    assert not set(rep["dropped_items"].tolist()) & set(kept.tolist())
    assert rep["n_clusters"] == len(o.clusters)


@pytest.mark.gpu
def test_device_linkage_matches_scipy_and_helper():
    from imageclust_b200 import _lib
    rng = np.random.default_rng(21)
    n = 600
    x = rng.standard_normal((n, 24)).astype(np.float32)
    mn, mx = _unconstrained(n)
    with clustering.Engine(0) as eng:
        eng.set_option("gram_mode", _lib.GRAM_EXACT_FP32)
        res = eng.cluster(x, mn, mx)
        z = eng.linkage()
        tr = eng.merge_trace()
        rep = clustering.clustering_report(n, res.clusters, res.stats)
    assert np.array_equal(z, clustering.trace_to_linkage(tr))
    want = linkage(x.astype(np.float64), "ward")[: len(z)]
    np.testing.assert_allclose(z[:, 2], want[:, 2], rtol=5e-5)
    assert np.array_equal(z[:, 3], want[:, 3])
    assert np.array_equal(np.sort(z[:, :2], axis=1), np.sort(want[:, :2], axis=1))
    assert len(rep["dropped_items"]) == 0 and rep["n_clusters"] == res.stats["n_out"]
