#!/usr/bin/env python
"""bench.py -- size-constrained Ward clustering on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config A|B|C|E] [--impl reference]

One "step" = one whole PerformClusteringWithConstraints pass (clustering.go:198-284)
over one synthetic Gaussian-mixture embedding matrix.  Default workload: BASELINE
config C, N=100,000 x 2048, minSize=20, maxSize=200 (the size the metric is quoted
on; its 40 GB distance matrix fits one B200).  Prints ONE JSON line (rank 0).

value    seconds per clustering with X already resident in HBM (ic_run_resident)
e2e      the same through the reference-facing call with HOST (pinned) buffers:
         H2D of X and D2H of the merge trace inside the timed region
roofline the dominant kernel (the persistent merge loop, HBM-bound by design,
         latency-bound in practice) + the other kernels under "kernels"
cpu_baseline  the CPU oracle (C restatement of the Go reference) on a bounded sample

With --gpus N > 1 (torchrun) the SAME clustering is row-block sharded over the N GPUs
(BASELINE config "N=100,000 x 2048 on 1 x B200 vs 8 x B200 row-block sharded"): rank r keeps
the rows of its slot block, the ranks' persistent kernels exchange their candidate pairs once per
iteration (~10 merges) through peer-mapped memory over NVLink (imageclust_b200/sharding.py).
Total work is fixed => "scaling": "strong".  --replicas runs N independent clusterings
instead (one per GPU, "weak").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def loop_traffic_gb(config):
    """dram__bytes_read.sum + dram__bytes_write.sum of the merge-loop launches of one clustering (GB), from the committed
    ncu --set full capture of this round's kernel (profiles/r02_loop_traffic.json); None if not captured for the config."""
    p = os.path.join(ROOT, "profiles", "r02_loop_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(config)
    return None

from imageclust_b200 import synth  # noqa: E402

METRIC = "ward_clustering_wall_time_s"
UNIT = "s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j["bf16_tflops"]),
                "bf16_tflops_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "source": "measured"}
    # /opt/skills/guides/B200_PROFILING.md fallback
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload(args):
    n, d, mn, mx = synth.CONFIGS[args.config]
    if args.n:
        n = args.n
    return n, d, mn, mx


def make_matrix(args, n, d, mn, mx, rank, out):
    # sharded run: every rank holds the same matrix; replicas: one matrix per rank
    seed = 20240 + ord(args.config) - ord("A") + (1000 * rank if args.replicas else 0)
    if args.config == "E":
        out[:] = synth.combined_features(n, 2048, d - 2048, mn, mx, seed=seed)
    else:
        synth.gaussian_mixture(n, d, mn, mx, seed=seed, out=out)
    return out


# ------------------------------------------------------------------------------------------
# CPU legs: the ONLY places this file executes anything under oracle/
# ------------------------------------------------------------------------------------------

def cpu_literal_sample(n_sample, d, mn, mx, seed=20241):
    """C restatement of the Go reference (oracle/ward_literal.c), single thread like the Go code."""
    from oracle import oracle as O
    x = synth.gaussian_mixture(n_sample, d, mn, mx, seed=seed)
    t0 = time.perf_counter()
    r = O.literal_cluster(x, mn, mx)
    dt = time.perf_counter() - t0
    assert r.ok
    return dt, r.n_merges


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm (C restatement; Go cannot be built here)
    on the host cores, each step a bounded sample of the workload."""
    if rank != 0:
        return
    n, d, mn, mx = workload(args)
    n_sample = min(n, args.ref_n)
    mn_s, mx_s = (mn, mx) if n_sample >= 4 * mx else (5, 20)
    for _ in range(min(args.warmup, 1)):
        cpu_literal_sample(min(n_sample, 600), d, mn_s, mx_s)
    times = []
    merges = 0
    for s in range(args.steps):
        dt, merges = cpu_literal_sample(n_sample, d, mn_s, mx_s, seed=20241 + s)
        times.append(dt)
    v = sum(times) / len(times)
    sample = (f"literal C restatement of clustering.go (not Go), 1 thread (the Go code is single threaded), "
              f"N={n_sample} x {d}, min/max {mn_s}/{mx_s}, {merges} merges per step; the full N={n} workload "
              f"scales ~ (N/{n_sample})^3 => ~{v * (n / n_sample) ** 3:.3g} s")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config {args.config}: N={n} x {d} min/max {mn}/{mx}; reference arm sample N={n_sample}"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(d, eng, budget_n=2000):
    """CPU numbers beside the GPU line, and the oracle used AS THE CHECKER of the same samples run through the GPU path
    (``like_for_like``: BASELINE config 1; ``parity_sample``: N = 6000, the size at which Lance-Williams values alone
    diverge from the reference's arithmetic)."""
    import torch
    from oracle import oracle as O
    out = {}
    dt, merges = cpu_literal_sample(budget_n, d, 5, 20)
    # strongest CPU comparator: same semantics, NN cache, all cores
    cores = os.cpu_count() or 1
    nf = 6000
    x = synth.gaussian_mixture(nf, d, 10, 50, seed=20242)
    t0 = time.perf_counter()
    r = O.fast_cluster(x, 10, 50, n_threads=cores)
    dtf = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": dt, "unit": UNIT, "cores": 1, "kind": "port",
                           "sample": f"literal C restatement of clustering.go (not Go), 1 thread, N={budget_n} x {d}, 5/20, {merges} merges",
                           "fast_oracle": {"value": dtf, "unit": UNIT, "cores": cores,
                                           "sample": f"oracle_fast (same results, NN cache, OpenMP) N={nf} x {d}, 10/50, {r.n_merges} merges"},
                           "host_cores": cores}
    # the GPU path (default tensor-core Gram) on the same N = 6000 sample, checked against that reference-arithmetic run
    xs = eng.pinned_empty((nf, d))
    xs[:] = x
    res = eng.cluster(xs, 10, 50)
    tr = eng.merge_trace()
    same = (len(tr.key_hi) == r.n_merges and np.array_equal(tr.key_hi, r.key_hi) and np.array_equal(tr.key_lo, r.key_lo))
    first = -1
    if not same:
        m = min(len(tr.key_hi), r.n_merges)
        neq = np.flatnonzero((tr.key_hi[:m] != r.key_hi[:m]) | (tr.key_lo[:m] != r.key_lo[:m]))
        first = int(neq[0]) if len(neq) else m
    from sklearn.metrics import adjusted_rand_score
    lab_a = np.full(nf, -1)
    lab_b = np.full(nf, -1)
    for cid, c in enumerate(res.clusters):
        lab_a[c] = cid
    for cid, c in enumerate(r.clusters):
        lab_b[c] = cid
    out["parity_sample"] = {"workload": f"N={nf} x {d}, min/max 10/50, default Gram, vs oracle in reference arithmetic (flags=0)",
                            "ari_vs_reference": float(adjusted_rand_score(lab_a, lab_b)), "trace_identical": bool(same),
                            "dist_bits_identical": bool(same and np.array_equal(tr.dist.view(np.uint32), r.dist.view(np.uint32))),
                            "first_divergence": first, "n_filter_viol": res.stats["n_filter_viol"],
                            "n_order_viol": res.stats["n_order_viol"]}
    # BASELINE config 1 (N = 1000 x 2048, 5/20) on both arms: the literal CPU restatement (1 thread) and the GPU path end to
    # end from host memory through ic_cluster_with_constraints
    n1, d1, mn1, mx1 = synth.CONFIGS["A"]
    x1 = synth.gaussian_mixture(n1, d1, mn1, mx1, seed=20240)
    t0 = time.perf_counter()
    lit = O.literal_cluster(x1, mn1, mx1)
    t_cpu = time.perf_counter() - t0
    xp = eng.pinned_empty((n1, d1))
    xp[:] = x1
    eng.cluster(xp, mn1, mx1)
    reps = 10
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r1 = eng.cluster(xp, mn1, mx1)
    torch.cuda.synchronize()
    t_gpu = (time.perf_counter() - t0) / reps
    tr1 = eng.merge_trace()
    out["like_for_like"] = {"workload": f"BASELINE config 1: N={n1} x {d1}, min/max {mn1}/{mx1}, {lit.n_merges} merges, same matrix on both arms",
                            "gpu_e2e_s": t_gpu, "cpu_literal_s": t_cpu, "cpu_cores": 1, "ratio": t_cpu / t_gpu,
                            "same_result": bool(len(tr1.key_hi) == lit.n_merges and np.array_equal(tr1.key_hi, lit.key_hi)
                                                and np.array_equal(tr1.key_lo, lit.key_lo)
                                                and np.array_equal(tr1.dist.view(np.uint32), lit.dist.view(np.uint32))
                                                and len(r1.clusters) == len(lit.clusters)
                                                and all(np.array_equal(a, b) for a, b in zip(r1.clusters, lit.clusters)))}
    return out


def measure_tensor_peaks():
    """TF32 and int8 dense matmul throughput of this GPU, measured the way MEASURED_PEAKS.json's bf16 figure was
    (torch.matmul / torch._int_mm 8192^3, best of 10): the denominators of K1's roofline lines."""
    import torch
    out = {}
    n = 8192
    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        best = 1e9
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["tf32_tflops"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
    except Exception as e:  # noqa: BLE001
        out["tf32_error"] = str(e)[:80]
    try:
        a = torch.randint(-64, 64, (n, n), device="cuda", dtype=torch.int8)
        b = torch.randint(-64, 64, (n, n), device="cuda", dtype=torch.int8)
        best = 1e9
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["int8_tops"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        del a, b
    except Exception as e:  # noqa: BLE001
        out["int8_error"] = str(e)[:80]
    torch.cuda.empty_cache()
    return out


def trace_sha(tr):
    import hashlib
    return hashlib.sha256(tr.key_hi.tobytes() + tr.key_lo.tobytes() + tr.dist.tobytes() + tr.size.tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C", choices=sorted(synth.CONFIGS))
    ap.add_argument("--n", type=int, default=0, help="override N (debug)")
    ap.add_argument("--ref-n", type=int, default=1500, help="sample size of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gram-mode", type=int, default=2, help="2: tcgen05 kind::i8 (default), 0: tcgen05 kind::tf32, 1: exact fp32 SIMT")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent clusterings instead of one sharded one")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from imageclust_b200 import clustering, sharding

    n, d, mn, mx = workload(args)
    peaks = load_peaks()
    eng = clustering.Engine(local_rank)
    eng.set_option("gram_mode", args.gram_mode)
    sharded = world > 1 and not args.replicas
    x_host = eng.pinned_empty((n, d))
    make_matrix(args, n, d, mn, mx, rank, x_host)
    if sharded:
        # one clustering over all ranks: the wrapper exchanges the CUDA-IPC handles once, everything
        # else (load / run_resident / cluster) is the same call on every rank
        runner = sharding.ShardedEngine(eng, rank, world)
    else:
        runner = eng

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- cold call: the first clustering of this process (allocations, first launches) through the e2e entry point ----
    barrier()
    t0 = time.perf_counter()
    runner.cluster(x_host, mn, mx)
    barrier()
    cold_s = max_over_ranks(time.perf_counter() - t0)

    # ---- resident leg: X in HBM before the timed region --------------------------------
    runner.load(x_host)
    stats = []
    for _ in range(args.warmup):
        runner.run_resident(mn, mx)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = runner.run_resident(mn, mx)
        stats.append(r.stats)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    sec_per_step = dt / args.steps
    sha = trace_sha(eng.merge_trace())  # every rank records the same trace; N = 1, 2, 4, 8 must print the same digest
    ranks_agree = True
    if world > 1:
        shas = [None] * world
        dist.all_gather_object(shas, sha)
        ranks_agree = len(set(shas)) == 1

    # ---- e2e leg: host buffers through the reference-facing call ---------------------------
    for _ in range(1):
        runner.cluster(x_host, mn, mx)
    barrier()
    t0 = time.perf_counter()
    e2e_stats = []
    for _ in range(args.steps):
        r = runner.cluster(x_host, mn, mx)
        e2e_stats.append(r.stats)
    barrier()
    e2e_dt = max_over_ranks(time.perf_counter() - t0)
    e2e_sec = e2e_dt / args.steps

    # ---- kernel microbenchmarks for the roofline legs (rank 0) ----------------------------
    kern = {}
    tpeaks = {}
    if rank == 0 and not sharded:
        eng.load(x_host)  # (releases nothing: same shape) -- the 40 GB matrix stays; the 8192^3 probes need < 1 GB
        tpeaks = measure_tensor_peaks()
        kern["gram_ms"] = eng.time_kernel({0: "gram", 1: "gram_exact", 2: "gram_i8"}[args.gram_mode],
                                          1 if args.gram_mode == 1 else 3)
        kern["nn_sweep_ms"] = eng.time_kernel("nn_sweep", 5)
    if world > 1:
        dist.barrier()

    if rank == 0:
        def avg(key, src=stats):
            return sum(s[key] for s in src) / len(src)

        merges = stats[-1]["n_merges"]
        batched = bool(stats[-1].get("loop_mode", 0))
        n_final = stats[-1]["n_final"]
        pairs = n * (n - 1) / 2
        ms_loop = avg("ms_loop")
        ms_gram = avg("ms_gram")
        ms_prep = avg("ms_prep")
        ms_nn = avg("ms_nn_init")
        # algorithmic bytes of the merge loop: 12*n per merge (read row a, read row b, write the new row), over the device time of
        # the loop kernel's launches (CUDA events around every launch, summed: ic_stats.ms_loop_kernel); ms_loop additionally
        # holds the horizon sweeps (refine.cu, near.cu), the compactions and the host round trips between the launches
        loop_bytes = 12.0 * sum(range(n_final + 1, n + 1)) if merges else 0.0
        ms_loop_kernel = avg("ms_loop_kernel") if stats[-1].get("ms_loop_kernel", 0) > 0 else ms_loop
        loop_gbs = loop_bytes / (ms_loop_kernel * 1e-3) / 1e9 if ms_loop_kernel > 0 else 0.0
        hbm = peaks["hbm_gbs"]
        # K1: 2*D flops per unordered pair; peak = TF32 dense = half the measured bf16 figure
        gram_flops = 2.0 * d * pairs
        # the time_kernel figure is an isolated launch: burst denominators.  Measured TF32 / int8 matmul peaks of this GPU
        # (torch 8192^3, this run) where available, else bf16 / 2 and bf16 x 2 of MEASURED_PEAKS
        tf32_peak = tpeaks.get("tf32_tflops") or peaks["bf16_tflops"] / 2.0
        int8_peak = tpeaks.get("int8_tops") or peaks["bf16_tflops"] * 2.0
        gram_tf = gram_flops / (kern["gram_ms"] * 1e-3) / 1e12 if kern.get("gram_ms") else 0.0
        sweep_gbs = 4.0 * pairs / (kern["nn_sweep_ms"] * 1e-3) / 1e9 if kern.get("nn_sweep_ms") else 0.0
        phases = {k: avg(k) for k in ("ms_prep", "ms_gram", "ms_nn_init", "ms_loop", "ms_d2h", "ms_host", "ms_total")}
        dominant = max(("ms_loop", "ms_gram", "ms_nn_init", "ms_prep"), key=lambda k: phases[k])
        line = {
            "metric": METRIC, "value": sec_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": False,
            # one clustering of fixed size, on 1 GPU or row-block sharded over N: total work does not grow with N
            "scaling": "weak" if (world > 1 and args.replicas) else "strong",
            "trace_sha": sha, "ranks_agree": ranks_agree, "cold_first_call_s": cold_s,
            "vs_baseline": None, "dtype": {0: "tf32+f32", 1: "f32", 2: "i8/i32 (22-bit fixed point) + f32"}[args.gram_mode], "data": "synthetic",
            "config": {"workload": f"config {args.config}: N={n} x {d} Gaussian-mixture fp32 embeddings, "
                                   f"minSize={mn}, maxSize={mx} -> {stats[-1]['n_target']} clusters, {merges} merges",
                       "parallelism": "1 GPU" if world == 1 else (
                           f"one clustering row-block sharded over {world} GPUs (peer-mapped rows; the ranks' persistent kernels "
                           f"exchange their candidate pairs once per iteration over NVLink)" if sharded else f"{world} independent clustering jobs (one per GPU)"),
                       "l2": "inputs larger than L2 (X %.0f MB, distance matrix %.1f GB)" % (4e-6 * n * d, stats[-1]["matrix_bytes"] / 1e9),
                       "gram": {0: "tcgen05 kind::tf32, exact fixed-point slice + residual (4 products)", 1: "exact fp32 SIMT", 2: "tcgen05 kind::i8, three int8 digits of a 22-bit fixed-point row, exact int32 accumulation (6 products)"}[args.gram_mode]},
            "merges_per_s": merges / (ms_loop * 1e-3) if ms_loop > 0 else None,
            "dist_matrix_gbs": 4.0 * pairs / (ms_gram * 1e-3) / 1e9 if ms_gram > 0 else None,
            "phases_ms": phases,
            "n_near_ties": stats[-1]["n_near_ties"], "n_rescans": stats[-1]["n_rescans"],
            # reference arithmetic: pairs re-evaluated as WardDistance of two fp32 centroids, and the
            # two guarantees that make the merge sequence the reference's (both must be 0)
            "reference_arithmetic": {k: stats[-1][k] for k in ("exact", "n_exact", "n_horizon_raises", "n_filter_viol", "n_order_viol",
                                                               "n_restarts", "n_compactions", "filter_max_err", "ms_refine", "ms_compact")},
            "exhausted": stats[-1]["exhausted"], "n_out": stats[-1]["n_out"],
            "e2e": {"value": e2e_sec, "unit": UNIT,
                    "h2d_bytes_per_step": e2e_stats[-1]["h2d_bytes"], "d2h_bytes_per_step": e2e_stats[-1]["d2h_bytes"],
                    "ms_h2d": sum(s["ms_h2d"] for s in e2e_stats) / len(e2e_stats)},
            "gpu_launches": int(sum(s["kernel_launches"] for s in stats)),
            "clocks": clocks,
            "loop": {"mode": "batched (merge_batch_kernel)" if batched else "one merge per iteration (merge_loop_kernel)",
                     "iterations": stats[-1]["n_iterations"],
                     "merges_per_iteration": merges / max(stats[-1]["n_iterations"], 1)},
            "roofline": {"kernel": "merge_batch_kernel (K3b, persistent, batched)" if batched else "merge_loop_kernel (K3, persistent)",
                         "bound": "hbm", "achieved": loop_gbs,
                         "peak": hbm, "unit": "GB/s", "frac": loop_gbs / hbm,
                         "launches_per_step": stats[-1].get("loop_launches"), "ms_kernel_per_step": ms_loop_kernel,
                         "ms_loop_per_step": ms_loop,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full at this exact
                         # workload (profiles/); GB per launch like `achieved`'s numerator (60 GB at config C)
                         "traffic": loop_traffic_gb(args.config) if (world == 1 and not args.n and batched) else None,
                         "traffic_unit": "GB per clustering, all launches of the loop kernel (ncu --set full, profiles/r02_summary.md)",
                         # what the launch actually moves over its measured time, against the same peak (information:
                         # the gap to `frac` is re-read rows, retired columns, partner lists, sector-granular gathers)
                         "traffic_frac": (loop_traffic_gb(args.config) / (ms_loop_kernel * 1e-3) / hbm
                                          if (world == 1 and not args.n and batched and loop_traffic_gb(args.config) and ms_loop_kernel > 0)
                                          else None),
                         "peak_source": peaks["source"] + " copy bandwidth",
                         "note": ("algorithmic bytes 12*n per merge over the summed device time of the kernel's launches (one per "
                                  "segment between horizon raises and compactions); an iteration takes every provably consecutive "
                                  "merge (dozens) in four grid-wide phases: partner lists, heads, Lance-Williams rows, exact "
                                  "re-evaluation; traffic = DRAM bytes of all launches of one clustering"
                                  if batched else
                                  "algorithmic bytes 12*n per merge; the loop is a chain of dependent merges bound by one "
                                  "mailbox exchange + one DRAM round trip per merge (merges_per_s), not by bandwidth"),
                         "dominant_phase": dominant},
            "kernels": None if sharded else {
                {0: "gram_tcgen05", 1: "gram_exact", 2: "gram_i8"}[args.gram_mode]: {
                    "bound": "tensor", "achieved": gram_tf, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": gram_tf / tf32_peak, "ms": kern.get("gram_ms"),
                    # what the pipe itself executes: 6 int8 products (4 TF32 products) per algorithmic product, against
                    # the int8 (TF32) dense peak = measured sustained bf16 x 2 (/ 2)
                    "pipe_frac": (gram_tf * 6.0 / int8_peak if args.gram_mode == 2 else
                                  gram_tf * 4.0 / tf32_peak if args.gram_mode == 0 else None),
                    "peak_source": ("torch.matmul TF32 8192^3 best of 12, this run" if tpeaks.get("tf32_tflops") else "MEASURED_PEAKS bf16 burst / 2"),
                    "pipe_peak": int8_peak if args.gram_mode == 2 else tf32_peak,
                    "pipe_peak_source": ("torch._int_mm int8 8192^3 best of 12, this run" if (args.gram_mode == 2 and tpeaks.get("int8_tops")) else
                                         "MEASURED_PEAKS bf16 burst x 2" if args.gram_mode == 2 else "as peak"),
                    "measured_peaks": tpeaks,
                    "note": "isolated launch (ic_time_kernel); frac = algorithmic flops (2*D per unordered pair) over the measured TF32 "
                            "matmul peak -- the fp32-accurate rate a plain GEMM would get; pipe_frac = what the tensor pipe executes "
                            "(6 kind::i8 MMAs per k-step = 6x the algorithmic flops; the tf32 path 4 kind::tf32 MMAs) over the measured "
                            "int8 (TF32) matmul peak"},
                "nn_sweep": {"bound": "hbm", "achieved": sweep_gbs, "peak": hbm, "unit": "GB/s", "frac": sweep_gbs / hbm,
                             "ms": kern.get("nn_sweep_ms"), "note": "algorithmic bytes 4 per pair (lower triangle read once)"},
            },
        }
        # the SAME bounded sample the reference arm times per step (--impl reference: N = --ref-n items of this config,
        # seed 20241), through the reference-facing call with host buffers: the like-for-like number beside that arm
        if world == 1 and not sharded:
            n_s = min(n, args.ref_n)
            mn_s, mx_s = (mn, mx) if n_s >= 4 * mx else (5, 20)
            xs = eng.pinned_empty((n_s, d))
            synth.gaussian_mixture(n_s, d, mn_s, mx_s, seed=20241, out=xs)
            eng.cluster(xs, mn_s, mx_s)
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                rs = eng.cluster(xs, mn_s, mx_s)
            torch.cuda.synchronize()
            line["reference_sample"] = {"workload": f"N={n_s} x {d}, min/max {mn_s}/{mx_s} (one step of --impl reference)",
                                        "value": (time.perf_counter() - t0) / reps, "unit": UNIT, "merges": rs.stats["n_merges"]}
        if not args.no_cpu_baseline and world == 1:
            line.update(cpu_baseline_leg(d, eng))
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)

    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
