"""Tiny profiling target: ONE resident clustering pass (every kernel of the path once).
usage: python scripts/profile_target.py [config] [n-override] [passes]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageclust_b200 import clustering, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "B"
n, d, mn, mx = synth.CONFIGS[cfg]
if len(sys.argv) > 2 and int(sys.argv[2]) > 0:
    n = int(sys.argv[2])
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
x = synth.gaussian_mixture(n, d, mn, mx, seed=20240 + ord(cfg) - ord("A"))
with clustering.Engine(0) as eng:
    eng.set_option("profile_loop", 1)
    if os.environ.get("IC_SCAN_EVERY"):
        eng.set_option("scan_every", int(os.environ["IC_SCAN_EVERY"]))
    mode = int(os.environ.get("IC_LOOP_MODE", "1"))
    eng.set_option("loop_mode", mode)
    if os.environ.get("IC_LOOP_DEBUG"):
        eng.set_option("loop_debug", int(os.environ["IC_LOOP_DEBUG"]))
    if os.environ.get("IC_LOOP_BLOCKS"):
        eng.set_option("loop_blocks", int(os.environ["IC_LOOP_BLOCKS"]))
    eng.load(x)
    for _ in range(passes):
        r = eng.run_resident(mn, mx)
    s = r.stats
    print(f"config {cfg} N={n} D={d}: merges={s['n_merges']} out={s['n_out']} prep {s['ms_prep']:.3f} gram {s['ms_gram']:.3f} "
          f"nn {s['ms_nn_init']:.3f} loop {s['ms_loop']:.3f} ms  rescans={s['n_rescans']} near_ties={s['n_near_ties']}")
    import hashlib
    tr = eng.merge_trace()
    print("trace_sha=" + hashlib.sha256(tr.key_hi.tobytes() + tr.key_lo.tobytes() + tr.dist.tobytes()).hexdigest()[:12])
    p = eng.loop_profile()
    m = max(p["merges"], 1)
    if mode == 1:  # batched loop: cycles of block 0 per phase
        it = max(p["iterations"], 1)
        print(f"batched loop: iterations={p['iterations']} merges/iteration={p['merges'] / it:.1f} cycles per iteration: "
              f"lists+rescans={p['publish'] / it:.0f} heads={p['exchange'] / it:.0f} select={p['update'] / it:.0f} apply={p['scan'] / it:.0f} "
              f"(rows+centroids={p['pub_argmin'] / it:.0f}, wait for the slowest block={p['exch_poll'] / it:.0f}, exact phase={p['exch_rank'] / it:.0f}) | "
              f"block 0 waits at the barriers: after rescans={p['pub_reduce'] / it:.0f} after heads={p['pub_fence'] / it:.0f}; "
              f"select = load+minima {p['fold'] / it:.0f} + theta {p['rescans'] / it:.0f} + filter {p['reserved'] / it:.0f} + conflicts/ranks; candidates below theta {p['bubbles'] / it:.1f}; "
              f"exact={s['exact']} n_exact={s['n_exact']} raises={s['n_horizon_raises']} compactions={s['n_compactions']} "
              f"refine+near {s['ms_refine']:.1f} ms compact {s['ms_compact']:.1f} ms filter_viol={s['n_filter_viol']} order_viol={s['n_order_viol']}")
        sys.exit(0)
    print("loop cycles per merge (block 0): " + " ".join(f"{k}={v / m:.0f}" for k, v in p.items() if k not in ("merges", "iterations", "rescans", "reserved", "bubbles"))
          + f" | iterations={p["iterations"]} rescans={p["rescans"]} bubbles={p["bubbles"]}")
    w = eng.loop_block_waits() / max(p["iterations"], 1)
    if len(w):
        order = np.argsort(w)
        print(f"exchange wait per iteration and block: min {w.min():.0f} (block {order[0]}) median {np.median(w):.0f} max {w.max():.0f} "
              f"(block {order[-1]}); five smallest: " + " ".join(f"{int(b)}:{w[b]:.0f}" for b in order[:5]))
