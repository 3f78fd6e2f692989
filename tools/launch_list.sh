#!/bin/bash
# Launch list of the default bench command: per-launch gpu__time_duration.sum of OUR kernels (names k_*), clocks untouched.
# Without the name filter the first 400 launches of the 1 GB / 100 M-pattern workload are torch's input-generation kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_target.csv \
   python bench.py --steps 2 --warmup 3 --no-compact --no-cpu-baseline --no-e2e --no-extract --no-gather-peak > gpurun_out/launches_target.log 2>&1
echo "launch list rc=$?"
