#!/bin/bash
# Round-2 ncu evidence: launch list of the bench command, DRAM bytes of every kernel of one clustering (config C), and
# --set full captures of the heaviest loop launch (first epoch) and of the Gram kernel.  Each ncu run follows a plain run
# of the same command that exited 0.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/bench_C_plain_r02.json 2> gpurun_out/bench_C_plain_r02.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/ncu_launches_r02.log 2>&1
echo "launch list exit $?"
CMD="python scripts/profile_target.py C 0 1"
$CMD > gpurun_out/plain_r02.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/dram_r02.csv $CMD > gpurun_out/ncu_dram_r02.log 2>&1
echo "dram pass exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'merge_batch' --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r02_loop $CMD > gpurun_out/ncu_full_loop_r02.log 2>&1
echo "full capture (loop) exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_i8|refine_eval|nn_sweep' -c 3 -f -o gpurun_out/prof_r02_kernels $CMD > gpurun_out/ncu_full_kernels_r02.log 2>&1
echo "full capture (gram, sweep, refine) exit $?"
cat gpurun_out/plain_r02.log; ls -la gpurun_out/*.ncu-rep
