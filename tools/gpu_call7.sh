#!/bin/bash
# round-2 GPU call 7: extraction from the text (tests + bench lines), e2e copy floor, all workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --tb=short -rf -x -k "extraction" > gpurun_out/r02_c7_pytest.log 2>&1
echo "pytest extraction rc=$?"; tail -25 gpurun_out/r02_c7_pytest.log
for wl in target_dna1g cfg5_bytes1g cfg4_multi; do
  timeout 900 python bench.py --steps 10 --workload $wl > gpurun_out/r02_c7_bench_$wl.json 2> gpurun_out/r02_c7_bench_$wl.err
  echo "bench $wl rc=$?"; tail -c 400 gpurun_out/r02_c7_bench_$wl.err
done
timeout 600 python bench.py --steps 10 --workload cfg3_rlfm --mode rich --no-cpu-baseline > gpurun_out/r02_c7_bench_cfg3_rlfm_rich.json 2> gpurun_out/r02_c7_bench_cfg3_rlfm_rich.err
echo "bench cfg3 rich rc=$?"; tail -c 400 gpurun_out/r02_c7_bench_cfg3_rlfm_rich.err
echo done
