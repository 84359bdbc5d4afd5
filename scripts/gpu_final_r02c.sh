#!/bin/bash
# Final bench lines + launch list of the final code (round 2, second session).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=f4
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
timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"
head -c 400 gpurun_out/${T}_bench_C.json
