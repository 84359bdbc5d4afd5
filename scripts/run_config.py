"""Run one BASELINE config on the GPU with options and print the stats + trace digest.
usage: python scripts/run_config.py C [opt=value ...] [--reps 2] [--profile]"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageclust_b200 import clustering, synth  # noqa: E402


def digest(tr):
    return hashlib.sha256(tr.key_hi.tobytes() + tr.key_lo.tobytes() + tr.dist.tobytes() + tr.size.tobytes()).hexdigest()[:16]


def main():
    cfg = sys.argv[1]
    reps = 2
    opts = []
    profile = False
    args = sys.argv[2:]
    i = 0
    while i < len(args):
        if args[i] == "--reps":
            reps = int(args[i + 1])
            i += 2
        elif args[i] == "--profile":
            profile = True
            i += 1
        else:
            opts.append(args[i])
            i += 1
    if cfg in synth.CONFIGS:
        n, d, mn, mx = synth.CONFIGS[cfg]
        seed = 20240 + "ABCDE".index(cfg)
    else:
        n, d, mn, mx = (int(v) for v in cfg.split(","))
        seed = 20241
    x = synth.combined_features(n, 2048, 100, mn, mx, seed=20244) if cfg == "E" else synth.gaussian_mixture(n, d, mn, mx, seed=seed)
    eng = clustering.Engine(0)
    for kv in opts:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    if profile:
        eng.set_option("profile_loop", 1)
    eng.load(x)
    for r in range(reps):
        res = eng.run_resident(mn, mx)
        tr = eng.merge_trace()
        st = res.stats
        keys = ("n_merges", "n_iterations", "exact", "n_horizon_raises", "n_exact", "n_filter_viol", "n_order_viol", "n_cut",
                "filter_max_err", "horizon", "ms_refine", "ms_prep", "ms_gram", "ms_nn_init", "ms_loop", "ms_total", "n_rescans",
                "exhausted", "n_out")
        out = dict(cfg=cfg, opts=opts, rep=r, digest=digest(tr), **{k: st[k] for k in keys if k in st})
        if profile:
            out["profile"] = eng.loop_profile()
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
