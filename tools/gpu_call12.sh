#!/bin/bash
# round-2 GPU call 12: FINAL bench lines, every workload at full BASELINE size with CPU baseline + parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_final_bench_target_dna1g.json 2> gpurun_out/r02_final_bench_target_dna1g.err
echo "bench default rc=$?"; tail -c 300 gpurun_out/r02_final_bench_target_dna1g.err
for wl in cfg5_bytes1g cfg4_multi cfg3_rlfm cfg2_dna100m cfg1_dna1m; do
  timeout 900 python bench.py --steps 10 --workload $wl > gpurun_out/r02_final_bench_$wl.json 2> gpurun_out/r02_final_bench_$wl.err
  echo "bench $wl rc=$?"; tail -c 300 gpurun_out/r02_final_bench_$wl.err
done
timeout 600 python bench.py --steps 10 --workload cfg3_rlfm --mode rich --no-cpu-baseline > gpurun_out/r02_final_bench_cfg3_rlfm_rich.json 2> gpurun_out/r02_final_bench_cfg3_rlfm_rich.err
echo "bench cfg3 rich rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_reference_arm.json 2> gpurun_out/r02_final_bench_reference_arm.err
echo "reference arm rc=$?"; tail -c 600 gpurun_out/r02_final_bench_reference_arm.json
echo done
