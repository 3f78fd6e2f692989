#!/bin/bash
# round-2 GPU call 17: FINAL sources (block-local second pass default, RLFM AUTO -> resident SA, >2^32 group test):
# full parity suite with durations, per-step ncu capture of the target for profiles/ncu_traffic.json, default bench line,
# launch list of the default bench command, ncu --set full of the fused query kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 780 python -m pytest tests -m gpu -q --timeout=600 --tb=short -rf --durations=30 > gpurun_out/r02_c17_pytest.log 2>&1
echo "pytest rc=$?"; tail -45 gpurun_out/r02_c17_pytest.log
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum"
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -o gpurun_out/r02_c17_step_target_rich -f python tools/prof_step.py --workload target_dna1g > gpurun_out/r02_c17_step_target_rich.log 2>&1
echo "ncu rich rc=$?"; tail -2 gpurun_out/r02_c17_step_target_rich.log
timeout 600 python bench.py > gpurun_out/r02_c17_bench_target_dna1g.json 2> gpurun_out/r02_c17_bench_target_dna1g.err
echo "bench default rc=$?"; tail -c 400 gpurun_out/r02_c17_bench_target_dna1g.err; head -c 600 gpurun_out/r02_c17_bench_target_dna1g.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c17_launches_target.csv \
   python bench.py --steps 2 --warmup 3 --no-compact --no-cpu-baseline --no-e2e --no-extract --no-gather-peak > gpurun_out/r02_c17_launches_target.log 2>&1
echo "launch list rc=$?"
B="--steps 3 --no-compact --no-cpu-baseline --no-e2e --no-gather-peak --no-extract --npat 20000000"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_query_fused" --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r02_c17_fused_defer -f python bench.py $B > gpurun_out/r02_c17_ncu_fused_defer.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/r02_c17*
echo done
