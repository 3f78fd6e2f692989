#!/bin/bash
# round-2 GPU call 11: FINAL sources -- full parity suite, smoke(), per-step ncu captures (target: rich and compact) for profiles/ncu_traffic.json,
# launch list of the default bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/r02_c11_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r02_c11_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c11_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r02_c11_smoke.log
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum"
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -o gpurun_out/r02_c11_step_target_rich -f python tools/prof_step.py --workload target_dna1g > gpurun_out/r02_c11_step_target_rich.log 2>&1
echo "ncu rich rc=$?"
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -o gpurun_out/r02_c11_step_target_compact -f python tools/prof_step.py --workload target_dna1g --mode compact > gpurun_out/r02_c11_step_target_compact.log 2>&1
echo "ncu compact rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c11_launches_target.csv \
   python bench.py --steps 2 --warmup 3 --no-compact --no-cpu-baseline --no-e2e --no-extract --no-gather-peak > gpurun_out/r02_c11_launches_target.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/r02_c11*
echo done
