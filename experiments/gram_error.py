"""Experiment (GPU): accuracy of the tensor-core Gram value itself (raw accumulator sum)
against a float64 Gram of the centred data, for the product subsets of the split
x = s1 + rt (bit 2: s1*s1, bits 0/1: rt*s1 / s1*rt, bit 4: rt*rt; bit 3: raw output)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageclust_b200 import _lib, clustering, synth


def main():
    eng = clustering.Engine(0)
    for n, d in [(512, 64), (512, 512), (768, 2048), (512, 2148)]:
        x = synth.gaussian_mixture(n, d, 5, 20, seed=1 + d)
        mean = (x.astype(np.float64).sum(0) / n).astype(np.float32)
        xc = (x - mean).astype(np.float64)
        gref = xc @ xc.T
        low = np.tril(np.ones((n, n), bool), -1)
        eng.load(x)
        for terms in (23, 7, 4):
            eng.set_option("gram_terms", terms | 8)
            eng.initial_distances(_lib.GRAM_TCGEN05_3XTF32)
            g = eng.read_matrix().astype(np.float64)
            err = (g - gref)[low]
            big = np.abs(gref[low]).max()
            print(f"n={n} d={d} terms={terms:2d}: rms err {np.sqrt((err ** 2).mean()):.3e} max|err| {np.abs(err).max():.3e} "
                  f"(max |g| {big:.1f}, ulp {big * 2 ** -23:.2e})")
    eng.set_option("gram_terms", 23)


if __name__ == "__main__":
    main()
