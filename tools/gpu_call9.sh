#!/bin/bash
# round-2 GPU call 9: k-mer table entries with text context (tests + A/B on the target), k = 4 table on the byte config, parity flags
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf -x -k "table_entries or rich_mode or scale_dna or packed or seed_and_verify" > gpurun_out/r02_c9_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02_c9_pytest.log
timeout 900 python bench.py --steps 10 > gpurun_out/r02_c9_bench_target_dna1g.json 2> gpurun_out/r02_c9_bench_target_dna1g.err
echo "bench target rc=$?"; tail -c 400 gpurun_out/r02_c9_bench_target_dna1g.err
Q="--no-cpu-baseline --no-compact --no-e2e --no-extract"
timeout 600 python bench.py --steps 10 $Q --option table_ctx=0 > gpurun_out/r02_c9_bench_target_ctx0.json 2> gpurun_out/r02_c9_bench_target_ctx0.err
echo "bench target ctx0 rc=$?"; tail -c 300 gpurun_out/r02_c9_bench_target_ctx0.err
timeout 900 python bench.py --steps 10 --workload cfg5_bytes1g $Q --option kmer_budget_mb=40000 > gpurun_out/r02_c9_bench_cfg5_k4.json 2> gpurun_out/r02_c9_bench_cfg5_k4.err
echo "bench cfg5 k4 rc=$?"; tail -c 300 gpurun_out/r02_c9_bench_cfg5_k4.err
timeout 900 python bench.py --steps 10 --workload cfg4_multi > gpurun_out/r02_c9_bench_cfg4_multi.json 2> gpurun_out/r02_c9_bench_cfg4_multi.err
echo "bench cfg4 rc=$?"; tail -c 300 gpurun_out/r02_c9_bench_cfg4_multi.err
timeout 900 python bench.py --steps 10 --workload cfg3_rlfm --no-compact > gpurun_out/r02_c9_bench_cfg3_rlfm.json 2> gpurun_out/r02_c9_bench_cfg3_rlfm.err
echo "bench cfg3 rc=$?"; tail -c 300 gpurun_out/r02_c9_bench_cfg3_rlfm.err
echo done
