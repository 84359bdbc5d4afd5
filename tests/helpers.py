"""Shared helpers for the test-suite (CPU side)."""
import glob
import hashlib
import os

import numpy as np

from imageclust_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    if "x" not in g:
        assert name == "cfgA_1000x2048"
        n, d, mn, mx = synth.CONFIGS["A"]
        g["x"] = synth.gaussian_mixture(n, d, mn, mx, seed=20240)
    assert hashlib.sha256(np.ascontiguousarray(g["x"]).tobytes()).hexdigest() == str(g["x_sha256"]), \
        "synthetic generator drifted from the golden input"
    return g


def golden_clusters(g):
    o, m = g["offsets"], g["members"]
    return [m[o[i]:o[i + 1]] for i in range(len(o) - 1)]


def same_clusters(a, b):
    return len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def labels_from_clusters(clusters, n):
    lab = np.full(n, -1, np.int64)
    for cid, c in enumerate(clusters):
        lab[np.asarray(c)] = cid
    return lab


def ari(clusters_a, clusters_b, n):
    """Adjusted Rand index of two (partial) partitions; dropped items (absent from
    the output map, clustering.go:268-271) form one extra 'dropped' class each side."""
    from sklearn.metrics import adjusted_rand_score
    return adjusted_rand_score(labels_from_clusters(clusters_a, n), labels_from_clusters(clusters_b, n))
