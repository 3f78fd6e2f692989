#!/bin/bash
# round-2 GPU call 4: parity with the device-built index (gpu_build.cu), target + cfg4 bench through it, per-step ncu traffic capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/r02_c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c4_pytest.log
tail -6 gpurun_out/r02_c4_pytest.log
for wl in target_dna1g cfg4_multi; do
  timeout 900 python bench.py --steps 10 --workload $wl > gpurun_out/r02_c4_bench_$wl.json 2> gpurun_out/r02_c4_bench_$wl.err
  echo "bench $wl rc=$?"; tail -c 300 gpurun_out/r02_c4_bench_$wl.err
done
FMX_HOST_BUILD=1 timeout 600 python bench.py --steps 3 --no-compact --no-cpu-baseline --no-e2e --no-gather-peak > gpurun_out/r02_c4_bench_target_hostbuild.json 2> gpurun_out/r02_c4_bench_target_hostbuild.err
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum"
for wl in target_dna1g; do
  timeout 900 ncu --metrics $M --clock-control none --profile-from-start off -o gpurun_out/r02_c4_step_$wl -f python tools/prof_step.py --workload $wl > gpurun_out/r02_c4_step_$wl.log 2>&1
  echo "ncu step $wl rc=$?"
done
ls -la gpurun_out/r02_c4*
echo done
