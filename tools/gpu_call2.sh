#!/bin/bash
# round-2 GPU call 2: parity incl. fused kernel / groups / scale tests, fused vs phased bench, ncu of the query kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/r02_c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c2_pytest.log
tail -5 gpurun_out/r02_c2_pytest.log
timeout 900 python bench.py --steps 10 > gpurun_out/r02_c2_bench_target.json 2> gpurun_out/r02_c2_bench_target.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02_c2_bench_target.err
B="--steps 3 --no-compact --no-cpu-baseline --no-e2e --no-gather-peak --npat 20000000"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_query_fused|k_emit_small" --launch-skip 8 --launch-count 2 \
   -o gpurun_out/r02_c2_fused -f python bench.py $B > gpurun_out/r02_c2_ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ph_seed|k_ph_verify|k_ph_steps" --launch-skip 12 --launch-count 4 \
   -o gpurun_out/r02_c2_phased -f python bench.py $B --option search_phased=1 > gpurun_out/r02_c2_ncu_phased.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
