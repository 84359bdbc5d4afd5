#!/bin/bash
# Final evidence of round 2 (second session): GPU tests, smoke, bench lines for configs C / A / B / E, then the ncu passes
# (launch list of the bench command, DRAM bytes per kernel of one clustering, --set full of the heaviest loop launch and of
# the Gram / sweep / refine kernels).  Every ncu run follows a plain run of the same command that exited 0.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=f2
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${T}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_C.json 2> gpurun_out/${T}_bench_C.err; echo "bench C exit $?"
for c in A B E; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; echo "bench $c exit $?"; done
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/${T}_bench_C_plain.json 2> gpurun_out/${T}_bench_C_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv $BENCH > gpurun_out/${T}_ncu_launches.log 2>&1
echo "launch list exit $?"
CMD="python scripts/profile_target.py C 0 1"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_dram.csv $CMD > gpurun_out/${T}_ncu_dram.log 2>&1
echo "dram pass exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'merge_batch' --launch-skip 1 --launch-count 1 -f -o gpurun_out/${T}_prof_loop $CMD > gpurun_out/${T}_ncu_full_loop.log 2>&1
echo "full capture (loop) exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_i8|refine_eval|compact_tiles' -c 3 -f -o gpurun_out/${T}_prof_kernels $CMD > gpurun_out/${T}_ncu_full_kernels.log 2>&1
echo "full capture (gram, refine, compaction) exit $?"
tail -n 2 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_plain.log | cut -c1-400; head -c 1500 gpurun_out/${T}_bench_C.json; ls -la gpurun_out/*.ncu-rep
