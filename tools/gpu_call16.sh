#!/bin/bash
# A/B: fused_defer = 2 (always) on config 4 (8-byte table entries) and config 5 (bytes, ragged)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-compact --no-e2e --no-extract"
for wl in cfg4_multi cfg5_bytes1g; do
 for o in 2 1; do
  timeout 600 python bench.py --steps 10 --workload $wl $Q --option fused_defer=$o > gpurun_out/r02_c16_bench_${wl}_defer$o.json 2> gpurun_out/r02_c16_bench_${wl}_defer$o.err
  echo "bench $wl defer=$o rc=$?"; tail -c 300 gpurun_out/r02_c16_bench_${wl}_defer$o.err
 done
done
echo done
