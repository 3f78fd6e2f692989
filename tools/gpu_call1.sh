#!/bin/bash
# round-2 GPU call 1: parity of the new paths, TLB-vs-DRAM microbenchmark, first target bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02_c1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/r02_c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c1_pytest.log
tail -5 gpurun_out/r02_c1_pytest.log
timeout 300 tools/tlb_window > gpurun_out/r02_tlb_window.jsonl 2> gpurun_out/r02_tlb_window.err
timeout 900 python bench.py --steps 10 > gpurun_out/r02_c1_bench_target.json 2> gpurun_out/r02_c1_bench_target.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r02_c1_bench_target.err
timeout 600 python bench.py --steps 10 --no-compact --no-cpu-baseline --no-e2e --option search_phased=0 > gpurun_out/r02_c1_bench_target_nophased.json 2> gpurun_out/r02_c1_bench_target_nophased.err
timeout 600 python bench.py --steps 10 --no-compact --no-cpu-baseline --no-e2e --option search_phased=0 --option locate_dense=0 > gpurun_out/r02_c1_bench_target_r1kernels.json 2> gpurun_out/r02_c1_bench_target_r1kernels.err
echo done
