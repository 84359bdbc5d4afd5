"""Multi-GPU check of the row-block sharded path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/shard_check.py [big]

Every rank clusters the same matrix as rank r of W shards; rank 0 runs the CPU oracle in REFERENCE arithmetic
(oracle.fast_cluster(flags=0), bit-identical to the literal restatement of clustering.go) and all ranks must hold
exactly that trace.  With IC_EXACT=0 the ranks keep Lance-Williams values only and the oracle's Lance-Williams mode
replays the gathered initial matrix instead."""
import hashlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageclust_b200 import _lib, clustering, sharding, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
big = len(sys.argv) > 1 and sys.argv[1] == "big"

eng = clustering.Engine(local)
eng.set_option("loop_mode", int(os.environ.get("IC_LOOP_MODE", "1")))  # 1: batched loop across the ranks, 0: one merge per iteration
exact_mode = int(os.environ.get("IC_EXACT", "1")) != 0 and int(os.environ.get("IC_LOOP_MODE", "1")) == 1
eng.set_option("exact", 1 if exact_mode else 0)
sh = sharding.ShardedEngine(eng, rank, world)
ok = True
cases = [(600, 48, 3, 10, _lib.GRAM_EXACT_FP32), (3001, 64, 4, 12, _lib.GRAM_TCGEN05_3XTF32),
         (5000, 256, 6, 8, _lib.GRAM_TCGEN05_I8),
         # 6000 exact duplicate pairs: every pair is a head at distance 0 and none conflicts, so each rank has ~3000
         # candidates below the stopper -- more than its region of the exchange box holds (kBatchXCand = 2048)
         (12000, 8, 1, 4, _lib.GRAM_EXACT_FP32)]
for n, d, mn, mx, mode in cases:
    if d == 8:
        rng = np.random.default_rng(11)
        base = (rng.standard_normal((n // 2, d)) * 50).astype(np.float32)
        x = np.concatenate([base, base])[rng.permutation(n)]
    else:
        x = synth.gaussian_mixture(n, d, mn, min(mx, 40), seed=n + d)
    eng.set_option("gram_mode", mode)
    sh.load(x)
    eng.initial_distances(mode, mx)
    m_own = torch.from_numpy(eng.read_matrix()).cuda()  # own rows, zeros elsewhere
    dist.all_reduce(m_own)
    eng.nn_init()
    eng.merge_loop(mn, mx)
    tr = eng.merge_trace()
    cl = eng.build_clusters(mn)
    digest = hashlib.sha256(tr.key_hi.tobytes() + tr.key_lo.tobytes() + tr.dist.tobytes() + tr.size.tobytes()).hexdigest()
    digests = [None] * world
    dist.all_gather_object(digests, digest)
    same = len(set(digests)) == 1
    if rank == 0:
        from oracle import oracle as O
        if exact_mode:
            o = O.fast_cluster(x, mn, mx, flags=0)
        else:
            m = np.tril(m_own.cpu().numpy(), -1)
            m = m + m.T
            o = O.fast_cluster(x, mn, mx, flags=O.FAST_EAGER | O.FAST_LW, init_matrix=m)
        exact = (len(tr.key_hi) == o.n_merges and np.array_equal(tr.key_hi, o.key_hi) and np.array_equal(tr.key_lo, o.key_lo)
                 and np.array_equal(tr.dist.view(np.uint32), o.dist.view(np.uint32))
                 and len(cl) == len(o.clusters) and all(np.array_equal(a, b) for a, b in zip(cl, o.clusters)))
        st = eng.stats()
        exact = exact and st["exact"] == (1 if exact_mode else 0) and st["n_filter_viol"] == 0 and st["n_order_viol"] == 0
        print(f"shard_check W={world} N={n} D={d} {mn}/{mx}: merges={st['n_merges']} ranks_agree={same} oracle_exact={exact} reference_arithmetic={exact_mode} "
              f"loop {st['ms_loop']:.2f} ms rescans={st['n_rescans']} iterations={st['n_iterations']} loop_mode={st['loop_mode']}", flush=True)
        ok = ok and same and exact
eng.set_option("gram_mode", _lib.GRAM_TCGEN05_I8)

if big:
    for cfg in sys.argv[2:] or ["B"]:
        n, d, mn, mx = synth.CONFIGS[cfg]
        x = eng.pinned_empty((n, d))
        synth.gaussian_mixture(n, d, mn, mx, seed=20240 + ord(cfg) - ord("A"), out=x)
        eng.set_option("profile_loop", 1)
        sh.load(x)
        for it in range(2):
            dist.barrier()
            t0 = time.perf_counter()
            r = sh.run_resident(mn, mx)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tr = eng.merge_trace()
            digest = hashlib.sha256(tr.key_hi.tobytes() + tr.key_lo.tobytes() + tr.dist.tobytes() + tr.size.tobytes()).hexdigest()
            digests = [None] * world
            dist.all_gather_object(digests, digest)
            s = r.stats
            if rank == 0:
                p = eng.loop_profile()
                mg = max(p["merges"], 1)
                print(f"config {cfg} W={world}: {dt:.3f} s  merges={s['n_merges']} out={s['n_out']} ranks_agree={len(set(digests)) == 1} "
                      f"prep {s['ms_prep']:.2f} gram {s['ms_gram']:.2f} nn {s['ms_nn_init']:.2f} loop {s['ms_loop']:.2f} ms "
                      f"rescans={s['n_rescans']} iterations={s['n_iterations']} loop_mode={s['loop_mode']} trace_sha={digest[:12]} exact={s['exact']} "
                      f"n_exact={s['n_exact']} filter_viol={s['n_filter_viol']} order_viol={s['n_order_viol']} refine {s['ms_refine']:.1f} ms | cycles/merge "
                      + " ".join(f"{k}={v / mg:.0f}" for k, v in p.items() if k in ("publish", "exchange", "update", "scan", "fold", "pub_fence", "pub_stores", "exch_poll", "exch_rank"))
                      + f" bubbles={p['bubbles']}", flush=True)
                ok = ok and len(set(digests)) == 1

flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
eng.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
