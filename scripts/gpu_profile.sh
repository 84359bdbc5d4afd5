#!/bin/bash
# ncu evidence for one resident pass: launch list (gpu__time_duration) + one --set full capture per kernel.
# usage: scripts/gpu_profile.sh [config] [n] [tag]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CFG="${1:-B}"; N="${2:-0}"; TAG="${3:-r01}"
CMD="python scripts/profile_target.py $CFG $N 1"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'merge_loop|gram_tcgen05|nn_sweep|split_kernel' -c 4 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
cat gpurun_out/plain_$TAG.log; tail -3 gpurun_out/ncu_full_$TAG.log; ls -la gpurun_out/
