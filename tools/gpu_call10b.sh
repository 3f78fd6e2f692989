#!/bin/bash
# A/B: ragged batches in order of length (config 5), + its test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=300 --tb=short -rf -x -k "ragged_batches or byte_alphabet_ragged" > gpurun_out/r02_c10b_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_c10b_pytest.log
Q="--no-cpu-baseline --no-compact --no-e2e --no-extract"
for o in 1 0; do
  timeout 600 python bench.py --steps 10 --workload cfg5_bytes1g $Q --option order_by_length=$o > gpurun_out/r02_c10b_bench_cfg5_order$o.json 2> gpurun_out/r02_c10b_bench_cfg5_order$o.err
  echo "bench cfg5 order=$o rc=$?"; tail -c 300 gpurun_out/r02_c10b_bench_cfg5_order$o.err
done
echo done
