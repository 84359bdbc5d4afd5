"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a
GPU, exports every symbol include/imageclust_b200.h declares, and its host logic
(CalculateOptimalClusters, clustering.go:168-186) matches the oracle.  No compute
entry point is called here."""
import os
import re

import numpy as np
import pytest

from imageclust_b200 import _lib, build, clustering

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "imageclust_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(ic_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"


def test_stats_struct_layout_matches_header():
    hdr = open(os.path.join(ROOT, "include", "imageclust_b200.h")).read()
    body = re.search(r"typedef struct ic_stats \{(.*?)\} ic_stats;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for ctype, name in re.findall(r"\b(int64_t|int32_t|float|double)\s+([a-z_0-9]+);", body):
        fields.append((name, ctype))
    got = [(n, {"c_long": "int64_t", "c_int": "int32_t", "c_float": "float", "c_double": "double"}[t.__name__])
           for n, t in _lib.Stats._fields_]
    assert fields == got


def test_optimal_clusters_matches_oracle(lib, oracle):
    cases = [(1000, 5, 20), (20000, 10, 50), (100000, 20, 200), (250000, 20, 200), (50000, 2, 8), (50000, 6, 8),
             (10, 5, 5), (3, 5, 20), (7, 4, 5), (0, 1, 5), (1, 1, 1), (17, 1, 1), (6, 6, 6)]
    rng = np.random.default_rng(0)
    cases += [tuple(int(v) for v in rng.integers(1, 60, 3)) for _ in range(300)]
    for n, mn, mx in cases:
        want, wrc = oracle.optimal_clusters(n, mn, mx)
        got, err = clustering.calculate_optimal_clusters(n, mn, mx)
        assert (err is None) == (wrc == 0), (n, mn, mx)
        if wrc == 0:
            assert got == want, (n, mn, mx)
    assert clustering.calculate_optimal_clusters(1000, 5, 20) == (125, None)
    n, err = clustering.calculate_optimal_clusters(3, 5, 20)
    assert n == 0 and "less than minimum cluster size" in err           # clustering.go:170
    n, err = clustering.calculate_optimal_clusters(7, 4, 5)
    assert n == 0 and "cannot satisfy cluster size constraints" in err  # clustering.go:176
    assert clustering.calculate_optimal_clusters(5, 0, 3)[1] is not None
    assert clustering.calculate_optimal_clusters(5, 1, 0)[1] is not None


def test_constraint_errors_return_nil_false_without_a_gpu(lib):
    # clustering.go:204-207: the constraint check comes before any device work
    ids = [f"img_{i}" for i in range(3)]
    x = np.zeros((3, 4), np.float32)
    assert clustering.perform_clustering_with_constraints(x, ids, 5, 20) == (None, False)
    x = np.zeros((7, 4), np.float32)
    assert clustering.perform_clustering_with_constraints(x, [f"i{i}" for i in range(7)], 4, 5) == (None, False)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the engine must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(clustering.EngineError):
        clustering.Engine(0)
    x = np.random.default_rng(0).standard_normal((16, 4)).astype(np.float32)
    with pytest.raises(clustering.EngineError):
        clustering.perform_clustering_with_constraints(x, [str(i) for i in range(16)], 2, 4)


def test_product_sources_do_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "imageclust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f
