"""compute-sanitizer target: one small golden and one small mixture through the batched loop (one rank and two virtual
ranks, with the horizon, near lists and a compaction) and through the one-merge-per-iteration loop."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageclust_b200 import clustering, synth  # noqa: E402

x = synth.gaussian_mixture(700, 64, 3, 10, seed=3)
with clustering.Engine(0) as eng:
    eng.set_option("compact_min", 64)
    ref = None
    for vr, mode in ((1, 1), (2, 1), (1, 0)):
        eng.set_option("virtual_ranks", vr)
        eng.set_option("loop_mode", mode)
        res = eng.cluster(x, 3, 10)
        st = res.stats
        print(f"virtual_ranks={vr} loop_mode={mode}: merges={st['n_merges']} iterations={st['n_iterations']} exact={st['exact']} "
              f"compactions={st['n_compactions']} n_exact={st['n_exact']} order_viol={st['n_order_viol']}", flush=True)
        if mode == 1:
            tr = eng.merge_trace()
            sig = (tr.key_hi.tobytes(), tr.key_lo.tobytes(), tr.dist.tobytes())
            assert ref is None or sig == ref, "virtual ranks changed the trace"
            ref = sig
print("sanitize target ok")
