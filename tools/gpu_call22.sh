#!/bin/bash
# round-2 GPU call 22 (the last 3 GPU-minutes): configs 4 and 5 on the final sources, device step only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
Q="--steps 10 --no-cpu-baseline --no-compact --no-e2e --no-extract"
timeout 70 python bench.py --workload cfg4_multi $Q > gpurun_out/r02_c22_bench_cfg4_multi_device_step.json 2> gpurun_out/r02_c22_cfg4.err
echo "cfg4 rc=$?"; head -c 250 gpurun_out/r02_c22_bench_cfg4_multi_device_step.json; echo
timeout 75 python bench.py --workload cfg5_bytes1g $Q > gpurun_out/r02_c22_bench_cfg5_bytes1g_device_step.json 2> gpurun_out/r02_c22_cfg5.err
echo "cfg5 rc=$?"; head -c 250 gpurun_out/r02_c22_bench_cfg5_bytes1g_device_step.json; echo
