#!/bin/bash
# One gpurun call: parity tests in isolated processes, smoke, a short bench.  Logs -> gpurun_out/.
# usage: scripts/gpu_round.sh [bench-config] [steps]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CFG="${1:-B}"
STEPS="${2:-2}"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
PT="python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600"
# group 1: everything that does not touch the tcgen05 kernel (a trap there poisons the CUDA context)
timeout 1500 $PT -k "not (tcgen05 or replays or config_a)" > gpurun_out/pytest_g1.log 2>&1
echo "g1 exit $?" >> gpurun_out/summary.txt
# group 2: tensor-core Gram kernel and the paths through it
timeout 1500 $PT -k "tcgen05 or replays or config_a" > gpurun_out/pytest_g2.log 2>&1
echo "g2 exit $?" >> gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 1200 python bench.py --config "$CFG" --steps "$STEPS" --warmup 3 > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err
echo "bench $CFG exit $?" >> gpurun_out/summary.txt
tail -n 3 gpurun_out/pytest_g1.log gpurun_out/pytest_g2.log gpurun_out/smoke.log
cat gpurun_out/summary.txt
head -c 3000 gpurun_out/bench_$CFG.json
