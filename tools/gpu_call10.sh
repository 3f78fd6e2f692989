#!/bin/bash
# round-2 GPU call 10: context-entry tests, ncu of the fused kernel with the 16-byte table, config 4 after pinning the occupancy
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=300 --tb=short -rf -x -k "table_entries" > gpurun_out/r02_c10_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02_c10_pytest.log
Q="--no-cpu-baseline --no-compact --no-e2e --no-extract"
timeout 600 python bench.py --steps 10 --workload cfg4_multi $Q > gpurun_out/r02_c10_bench_cfg4_multi.json 2> gpurun_out/r02_c10_bench_cfg4_multi.err
echo "bench cfg4 rc=$?"; tail -c 300 gpurun_out/r02_c10_bench_cfg4_multi.err
timeout 600 python bench.py --steps 10 $Q > gpurun_out/r02_c10_bench_target.json 2> gpurun_out/r02_c10_bench_target.err
echo "bench target rc=$?"; tail -c 300 gpurun_out/r02_c10_bench_target.err
B="--steps 3 --no-compact --no-cpu-baseline --no-e2e --no-gather-peak --no-extract --npat 20000000"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_query_fused" --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r02_c10_fused_ctx -f python bench.py $B > gpurun_out/r02_c10_ncu_fused_ctx.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_c10_fused_ctx.ncu-rep
echo done
