#!/bin/bash
# A/B: fused kernel with the block-local second pass (fused_defer), tests of the refactored head / tail
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --tb=short -rf -x -k "table_entries or rich_mode or packed or seed_and_verify or ragged" > gpurun_out/r02_c15_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_c15_pytest.log
Q="--no-cpu-baseline --no-compact --no-e2e --no-extract"
for o in 1 0; do
  timeout 600 python bench.py --steps 10 $Q --option fused_defer=$o > gpurun_out/r02_c15_bench_target_defer$o.json 2> gpurun_out/r02_c15_bench_target_defer$o.err
  echo "bench target defer=$o rc=$?"; tail -c 300 gpurun_out/r02_c15_bench_target_defer$o.err
done
echo done
