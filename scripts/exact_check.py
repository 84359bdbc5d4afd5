"""GPU check: device merge trace vs the CPU oracle in REFERENCE arithmetic (oracle.fast_cluster(flags=0), bit-identical
to the literal restatement of clustering.go).  Usage: python scripts/exact_check.py N [N ...] [--gram 1|2] [--min 10 --max 50]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageclust_b200 import clustering, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def compare(tr, r0):
    m = min(len(tr.key_hi), r0.n_merges)
    same = (tr.key_hi[:m] == r0.key_hi[:m]) & (tr.key_lo[:m] == r0.key_lo[:m])
    bits = tr.dist[:m].view(np.uint32) == r0.dist[:m].view(np.uint32)
    first = int(np.argmin(same)) if not same.all() else -1
    return dict(n_dev=len(tr.key_hi), n_ref=r0.n_merges, pairs_equal=bool(same.all() and len(tr.key_hi) == r0.n_merges),
                dist_bits_equal=bool(bits.all()), first_divergence=first, n_diverging=int((~same).sum()),
                n_dist_bits_differ=int((~bits).sum()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sizes", nargs="+", type=int)
    ap.add_argument("--gram", type=int, nargs="+", default=[1, 2])
    ap.add_argument("--min", type=int, default=10)
    ap.add_argument("--max", type=int, default=50)
    ap.add_argument("--d", type=int, default=2048)
    ap.add_argument("--seed", type=int, default=20241)
    ap.add_argument("--opt", nargs="*", default=[])
    a = ap.parse_args()
    eng = clustering.Engine(0)
    for kv in a.opt:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    for n in a.sizes:
        x = synth.gaussian_mixture(n, a.d, a.min, a.max, seed=a.seed)
        t0 = time.time()
        r0 = O.fast_cluster(x, a.min, a.max, flags=0)
        t_ref = time.time() - t0
        for g in a.gram:
            eng.set_option("gram_mode", g)
            t0 = time.time()
            res = eng.cluster(x, a.min, a.max)
            t_dev = time.time() - t0
            tr = eng.merge_trace()
            c = compare(tr, r0)
            st = res.stats
            same_map = len(res.clusters) == len(r0.clusters) and all(np.array_equal(p, q) for p, q in zip(res.clusters, r0.clusters))
            print(json.dumps(dict(n=n, gram=g, t_ref_s=round(t_ref, 2), t_dev_s=round(t_dev, 3), same_map=same_map, **c,
                                  stats={k: st[k] for k in ("exact", "n_merges", "n_iterations", "n_horizon_raises", "n_exact",
                                                            "n_filter_viol", "n_order_viol", "n_cut", "filter_max_err", "horizon",
                                                            "ms_refine", "ms_loop", "ms_gram", "ms_nn_init", "n_rescans")})), flush=True)


if __name__ == "__main__":
    main()
