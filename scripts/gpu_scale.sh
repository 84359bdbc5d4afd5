#!/bin/bash
# One multi-GPU gpurun call: bench.py sharded over N GPUs (+ optionally the N=250k config D check).
# usage: scripts/gpu_scale.sh N [steps] [warmup] [D]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N="${1:-2}"; STEPS="${2:-2}"; WARM="${3:-3}"; DOD="${4:-}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps $STEPS --warmup $WARM > gpurun_out/bench_C_g$N.json 2> gpurun_out/bench_C_g$N.err
echo "bench C x$N rc $?"; head -c 700 gpurun_out/bench_C_g$N.json; echo
if [ -n "$DOD" ]; then
  timeout 1500 $TR --master-port 29522 scripts/shard_check.py big D > gpurun_out/shard_D_g$N.log 2>&1
  echo "shard D x$N rc $?"; grep "shard_check\|config" gpurun_out/shard_D_g$N.log
fi
