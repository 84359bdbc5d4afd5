#!/bin/bash
# compute-sanitizer passes over both loop kernels on a small problem (racecheck: shared-memory hazards; memcheck: accesses)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 python scripts/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain exit $?"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_target.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_target.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?"
tail -n 4 gpurun_out/sanitize_plain.log gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log
