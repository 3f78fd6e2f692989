#!/bin/bash
# round-2 GPU call 8: full parity suite with the device-built SYM / RLFM blobs, build times of cfg3 / cfg5, memcheck of the new kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/r02_c8_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_c8_pytest.log
timeout 900 python bench.py --steps 10 --workload cfg5_bytes1g --no-cpu-baseline --no-compact > gpurun_out/r02_c8_bench_cfg5_bytes1g.json 2> gpurun_out/r02_c8_bench_cfg5_bytes1g.err
echo "bench cfg5 rc=$?"; tail -c 300 gpurun_out/r02_c8_bench_cfg5_bytes1g.err
timeout 900 python bench.py --steps 10 --workload cfg3_rlfm > gpurun_out/r02_c8_bench_cfg3_rlfm.json 2> gpurun_out/r02_c8_bench_cfg3_rlfm.err
echo "bench cfg3 rc=$?"; tail -c 300 gpurun_out/r02_c8_bench_cfg3_rlfm.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --timeout=800 -x \
   -k "gpu_built or extraction_from or wide_character_texts" > gpurun_out/r02_c8_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/r02_c8_memcheck.log
echo done
