"""Mint the golden vectors under tests/golden/ from the literal CPU oracle.

The reference (Go) cannot run in this image and ships no fixtures, so goldens are
produced by ``oracle/ward_literal.c`` (cross-checked bit for bit against
``oracle/numpy_literal.py`` in tests/test_oracle_literal.py).  Re-run with
``python tests/golden/make_golden.py``; the files are small and committed.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from imageclust_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def cases():
    rng = np.random.default_rng(7)
    yield "line6", np.array([[0], [1], [3], [6], [10], [15]], np.float32), 1, 3, True
    x = rng.standard_normal((8, 2)).astype(np.float32)
    yield "n8_d2", x, 2, 3, True
    x = rng.standard_normal((12, 2)).astype(np.float32)
    x[5] = x[2]; x[9] = x[2]; x[7] = x[1]  # exact ties at d = 0
    yield "ties12", x, 1, 4, True
    x = np.round(rng.standard_normal((24, 3)) * 2).astype(np.float32)  # integer grid: many equal distances
    yield "grid24", x, 2, 4, True
    yield "gm64", synth.gaussian_mixture(64, 32, 3, 8, seed=11), 3, 8, True
    yield "gm64_exhaust", synth.gaussian_mixture(64, 16, 6, 8, seed=12), 6, 8, True
    yield "gm256", synth.gaussian_mixture(256, 64, 4, 12, seed=13), 4, 12, True
    yield "gm256_relu", synth.gaussian_mixture(256, 64, 2, 6, seed=14, relu_like=True), 2, 6, True
    yield "cfgE_small", synth.combined_features(200, 64, 20, 2, 8, seed=15), 2, 8, True
    # BASELINE config A: X is regenerated from the seed (8 MB), its hash is stored
    n, d, mn, mx = synth.CONFIGS["A"]
    yield "cfgA_1000x2048", synth.gaussian_mixture(n, d, mn, mx, seed=20240), mn, mx, False


def main():
    for name, x, mn, mx, store_x in cases():
        r = O.literal_cluster(x, mn, mx, want_matrices=store_x)
        assert r.ok, name
        offs = np.cumsum([0] + [len(c) for c in r.clusters]).astype(np.int32)
        memb = np.concatenate(r.clusters).astype(np.int32) if r.clusters else np.zeros(0, np.int32)
        out = dict(min_size=mn, max_size=mx, n=x.shape[0], d=x.shape[1],
                   x_sha256=hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest(),
                   key_hi=r.key_hi, key_lo=r.key_lo, pos_i=r.pos_i, pos_j=r.pos_j, dist=r.dist,
                   size=r.size, offsets=offs, members=memb, n_target=r.n_target,
                   n_rejections=r.n_rejections, exhausted=int(r.exhausted), n_final=r.n_final)
        if store_x:
            out.update(x=x, init_matrix=r.init_matrix, final_matrix=r.final_matrix, final_keys=r.final_keys)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: N={x.shape[0]} D={x.shape[1]} {mn}/{mx} merges={r.n_merges} rej={r.n_rejections} "
              f"exhausted={r.exhausted} out={len(r.clusters)}")


if __name__ == "__main__":
    main()
