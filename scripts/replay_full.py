"""Full-size comparison of the GPU path with the CPU oracle in REFERENCE arithmetic (oracle.fast_cluster(flags=0): the
reference's centroid recompute, sequential fp32, bit-identical to the literal restatement of clustering.go; NN cache and
OpenMP only change the cost).  Usage: python scripts/replay_full.py E C [--min-size-e 2]
Prints one JSON line per config: merges, trace digests of both sides, first divergence, ARI, wall times."""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageclust_b200 import clustering, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def digest(key_hi, key_lo, dist, size):
    return hashlib.sha256(np.ascontiguousarray(key_hi).tobytes() + np.ascontiguousarray(key_lo).tobytes() +
                          np.ascontiguousarray(dist).tobytes() + np.ascontiguousarray(size).tobytes()).hexdigest()


def main():
    eng = clustering.Engine(0)
    for spec in sys.argv[1:]:
        parts = spec.split(":")  # config[:minSize[:oracle flags]]
        cfg = parts[0]
        n, d, mn, mx = synth.CONFIGS[cfg]
        if len(parts) > 1 and parts[1]:
            mn = int(parts[1])
        # flags 1 = FAST_EAGER: inadmissible pairs masked when written instead of rejected lazily -- same merge sequence
        # (tests/test_oracle_fast.py), without the reference's one-rescan-per-rejection cost (millions at min 6 / max 8)
        oflags = int(parts[2]) if len(parts) > 2 else 0
        seed = 20240 + "ABCDE".index(cfg)
        x = synth.combined_features(n, 2048, d - 2048, 2, 8, seed=seed) if cfg == "E" else synth.gaussian_mixture(n, d, mn, mx, seed=seed)
        t0 = time.time()
        res = eng.cluster(x, mn, mx)
        t_gpu = time.time() - t0
        tr = eng.merge_trace()
        st = res.stats
        t0 = time.time()
        o = O.fast_cluster(x, mn, mx, flags=oflags)
        t_cpu = time.time() - t0
        m = min(len(tr.key_hi), o.n_merges)
        neq = np.flatnonzero((tr.key_hi[:m] != o.key_hi[:m]) | (tr.key_lo[:m] != o.key_lo[:m]))
        same_map = len(res.clusters) == len(o.clusters) and all(np.array_equal(a, b) for a, b in zip(res.clusters, o.clusters))
        from sklearn.metrics import adjusted_rand_score
        la, lb = np.full(n, -1), np.full(n, -1)
        for cid, c in enumerate(res.clusters):
            la[c] = cid
        for cid, c in enumerate(o.clusters):
            lb[c] = cid
        print(json.dumps(dict(
            config=cfg, n=n, d=d, min_size=mn, max_size=mx, oracle_flags=oflags, merges_gpu=len(tr.key_hi), merges_oracle=o.n_merges,
            exhausted_gpu=bool(st["exhausted"]), exhausted_oracle=o.exhausted,
            digest_gpu=digest(tr.key_hi, tr.key_lo, tr.dist, tr.size), digest_oracle=digest(o.key_hi, o.key_lo, o.dist, o.size),
            first_divergence=int(neq[0]) if len(neq) else -1, dist_bits_identical=bool(len(neq) == 0 and len(tr.key_hi) == o.n_merges and np.array_equal(tr.dist.view(np.uint32), o.dist.view(np.uint32))),
            same_cluster_lists=bool(same_map), ari=float(adjusted_rand_score(la, lb)),
            gpu_wall_s=round(t_gpu, 3), oracle_wall_s=round(t_cpu, 1), oracle_threads=os.cpu_count(),
            n_exact=st["n_exact"], n_filter_viol=st["n_filter_viol"], n_order_viol=st["n_order_viol"], n_restarts=st["n_restarts"])), flush=True)


if __name__ == "__main__":
    main()
