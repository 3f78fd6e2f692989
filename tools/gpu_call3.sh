#!/bin/bash
# round-2 GPU call 3: parity after the word-wise compare / k-mer index, all workloads through the new bench, ncu of the fused kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/r02_c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c3_pytest.log
tail -4 gpurun_out/r02_c3_pytest.log
for wl in target_dna1g cfg5_bytes1g cfg2_dna100m cfg3_rlfm cfg4_multi cfg1_dna1m; do
  timeout 900 python bench.py --steps 10 --workload $wl > gpurun_out/r02_c3_bench_$wl.json 2> gpurun_out/r02_c3_bench_$wl.err
  echo "bench $wl rc=$?"; tail -c 300 gpurun_out/r02_c3_bench_$wl.err
done
timeout 600 python bench.py --steps 10 --workload cfg3_rlfm --mode rich --no-cpu-baseline > gpurun_out/r02_c3_bench_cfg3_rlfm_rich.json 2> gpurun_out/r02_c3_bench_cfg3_rlfm_rich.err
B="--steps 3 --no-compact --no-cpu-baseline --no-e2e --no-gather-peak --npat 20000000"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_query_fused" --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r02_c3_fused -f python bench.py $B > gpurun_out/r02_c3_ncu_fused.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_query_fused" --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r02_c3_fused_cfg5 -f python bench.py $B --workload cfg5_bytes1g > gpurun_out/r02_c3_ncu_fused_cfg5.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
