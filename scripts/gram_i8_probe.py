"""K1 (int8) experiment: time the kernel with parts switched off (wrong results on purpose).
usage: python scripts/gram_i8_probe.py [config]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageclust_b200 import clustering, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C"
n, d, mn, mx = synth.CONFIGS[cfg]
x = synth.gaussian_mixture(n, d, mn, mx, seed=20240 + ord(cfg) - ord("A"))
names = {0: "full kernel", 1: "no TMA loads", 2: "no MMAs", 4: "no epilogue stores", 6: "TMA loads only", 5: "MMAs only", 3: "epilogue only"}
with clustering.Engine(0) as eng:
    eng.load(x)
    for dbg in (0, 4, 1, 2, 6, 5, 3):
        eng.set_option("gram_debug", dbg)
        ms = eng.time_kernel("gram_i8", 3)
        print(f"gram_i8 config {cfg} debug={dbg} ({names[dbg]}): {ms:.2f} ms")
    eng.set_option("gram_debug", 0)
