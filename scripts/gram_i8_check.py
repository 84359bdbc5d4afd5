"""K1 int8 path: accuracy against the reference arithmetic and speed against the tf32 path (one GPU)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageclust_b200 import _lib, clustering, synth
from oracle import oracle as O

with clustering.Engine(0) as eng:
    for n, d, relu in [(256, 64, False), (700, 2048, False), (700, 2048, True), (333, 2148, False), (1000, 96, False)]:
        x = synth.gaussian_mixture(n, d, 5, 20, seed=100 + n + d, relu_like=relu)
        ref = O.initial_matrix(x)
        off = ~np.eye(n, dtype=bool)
        out = []
        for mode in (_lib.GRAM_TCGEN05_3XTF32, _lib.GRAM_TCGEN05_I8):
            eng.load(x)
            eng.initial_distances(mode)
            m = eng.read_matrix()
            out.append(float((np.abs(m[off] - ref[off]) / ref[off]).max()))
        print(f"n={n} d={d} relu={relu}: max rel err tf32 {out[0]:.2e}  i8 {out[1]:.2e}", flush=True)
    for cfg in sys.argv[1:] or ["B"]:
        n, d, mn, mx = synth.CONFIGS[cfg]
        x = synth.gaussian_mixture(n, d, mn, mx, seed=20240)
        eng.load(x)
        t_tf = eng.time_kernel("gram", 3)
        t_i8 = eng.time_kernel("gram_i8", 3)
        fl = 2.0 * d * n * (n - 1) / 2
        print(f"config {cfg}: gram tf32 {t_tf:.3f} ms ({fl / t_tf / 1e9:.1f} TFLOP/s alg)  i8 {t_i8:.3f} ms ({fl / t_i8 / 1e9:.1f} TFLOP/s alg)", flush=True)
