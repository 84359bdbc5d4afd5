#!/bin/bash
# End-of-round ncu evidence (batched merge loop): launch list of the bench command + one --set full capture of the
# path's kernels at config C.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/bench_C_plain_r01d.json 2> gpurun_out/bench_C_plain_r01d.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01d.csv $BENCH > gpurun_out/ncu_launches_r01d.log 2>&1
echo "launch list exit $?"
CMD="python scripts/profile_target.py C 0 1"
$CMD > gpurun_out/plain_r01d.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'merge_batch|gram_i8|nn_sweep' -c 3 -f -o gpurun_out/prof_r01d $CMD > gpurun_out/ncu_full_r01d.log 2>&1
echo "full capture exit $?"
cat gpurun_out/plain_r01d.log; tail -n 3 gpurun_out/ncu_full_r01d.log
