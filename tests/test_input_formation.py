"""Input formation: GenerateLabelVector + CombineEmbeddings
(/root/reference/internal/embeddings/embeddings.go:166-183, called at internal/workflow/workflow.go:167-168).
CPU: the oracle restatement on hand-checked cases.  GPU: ic_load_combined forms the same matrix in HBM, bit for bit,
and the clustering of it equals the clustering of the host-combined matrix."""
import numpy as np
import pytest

from imageclust_b200 import clustering, synth


def _case(seed=3, n=37, d_img=24, n_labels=11):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((n, d_img)).astype(np.float32)
    names = [f"label{j}" for j in range(n_labels)]
    label_set = {name: j for j, name in enumerate(names)}
    item_labels = []
    for i in range(n):
        k = int(rng.integers(0, 6))
        labs = [names[j] for j in rng.integers(0, n_labels, size=k)]  # duplicates on purpose
        if i % 5 == 0:
            labs.append("NotInTheSet")  # embeddings.go:169: ignored
        item_labels.append(labs)
    item_labels[1] = []  # no labels at all
    return img, item_labels, label_set


def _host_combined(oracle, img, item_labels, label_set):
    return np.stack([oracle.combine_embeddings(img[i], oracle.generate_label_vector(item_labels[i], label_set))
                     for i in range(len(img))])


def test_oracle_label_vector_and_combine(oracle):
    ls = {"Shoe": 0, "Red": 1, "Bag": 2}
    assert oracle.generate_label_vector(["Red", "Red", "Hat"], ls).tolist() == [0.0, 1.0, 0.0]
    assert oracle.generate_label_vector([], ls).tolist() == [0.0, 0.0, 0.0]
    assert oracle.generate_label_vector(["Bag", "Shoe"], {}).tolist() == []
    c = oracle.combine_embeddings(np.array([0.5, -2.0], np.float32), np.array([1.0, 0.0, 1.0], np.float32))
    assert c.dtype == np.float32 and c.tolist() == [0.5, -2.0, 1.0, 0.0, 1.0]
    img, labs, ls = _case()
    x = _host_combined(oracle, img, labs, ls)
    assert x.shape == (37, 24 + 11) and np.array_equal(x[:, :24], img)
    assert set(np.unique(x[:, 24:]).tolist()) <= {0.0, 1.0} and np.all(x[1, 24:] == 0)


@pytest.mark.gpu
def test_device_input_formation_is_bit_identical(oracle):
    img, labs, ls = _case()
    want = _host_combined(oracle, img, labs, ls)
    with clustering.Engine(0) as eng:
        eng.load_combined(img, labs, ls)
        got = eng.read_x()
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        # strided image block, empty label set, no items with labels
        wide = np.zeros((37, 40), np.float32)
        wide[:, :24] = img
        eng.load_combined(wide[:, :24], labs, ls)
        assert np.array_equal(eng.read_x(), want)
        eng.load_combined(img, [[] for _ in labs], {})
        assert np.array_equal(eng.read_x(), img)


@pytest.mark.gpu
def test_clustering_of_device_formed_matrix_equals_host_formed(oracle):
    n, d_img, n_labels = 600, 64, 20
    img = synth.gaussian_mixture(n, d_img, 2, 8, seed=11)
    rng = np.random.default_rng(12)
    names = [f"l{j}" for j in range(n_labels)]
    ls = {name: j for j, name in enumerate(names)}
    labs = [[names[j] for j in rng.choice(n_labels, size=int(rng.integers(0, 5)), replace=False)] for _ in range(n)]
    x = _host_combined(oracle, img, labs, ls)
    with clustering.Engine(0) as eng:
        a = eng.cluster(x, 2, 8)
        tr_a = eng.merge_trace()
        eng.load_combined(img, labs, ls)
        b = eng.run_resident(2, 8)
        tr_b = eng.merge_trace()
    assert np.array_equal(tr_a.key_hi, tr_b.key_hi) and np.array_equal(tr_a.key_lo, tr_b.key_lo)
    assert np.array_equal(tr_a.dist.view(np.uint32), tr_b.dist.view(np.uint32))
    assert len(a.clusters) == len(b.clusters) and all(np.array_equal(p, q) for p, q in zip(a.clusters, b.clusters))
