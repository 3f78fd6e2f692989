#!/bin/bash
# round-2 GPU call 19: FINAL sources (hash 15a4e14acfbb2704; the call-18 A/B kept the default instance): per-step ncu captures
# (rich + compact) summarised ON THE BOX into ncu_traffic.json (the .ncu-rep files stay there: gpurun_out is capped at 64 MiB),
# default bench line carrying the capture, launch list, ncu --set full summary, the other workloads' lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import bench; print('source hash', bench.source_hash())"
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum"
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -o /tmp/step_rich -f python tools/prof_step.py --workload target_dna1g > gpurun_out/r02_c19_step_target_rich.log 2>&1
echo "ncu rich rc=$?"; grep "^{" gpurun_out/r02_c19_step_target_rich.log
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -o /tmp/step_compact -f python tools/prof_step.py --workload target_dna1g --mode compact > gpurun_out/r02_c19_step_target_compact.log 2>&1
echo "ncu compact rc=$?"; grep "^{" gpurun_out/r02_c19_step_target_compact.log
python tools/ncu_traffic.py /tmp/step_rich.ncu-rep:gpurun_out/r02_c19_step_target_rich.log /tmp/step_compact.ncu-rep:gpurun_out/r02_c19_step_target_compact.log
cp profiles/ncu_traffic.json gpurun_out/r02_c19_ncu_traffic.json
ncu -i /tmp/step_rich.ncu-rep --page raw --csv > gpurun_out/r02_c19_step_target_rich_raw.csv 2>/dev/null
timeout 600 python bench.py > gpurun_out/r02_c19_bench_target_dna1g.json 2> gpurun_out/r02_c19_bench_target_dna1g.err
echo "bench default rc=$?"; tail -c 400 gpurun_out/r02_c19_bench_target_dna1g.err; head -c 300 gpurun_out/r02_c19_bench_target_dna1g.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c19_launches_target.csv \
   python bench.py --steps 2 --warmup 3 --no-compact --no-cpu-baseline --no-e2e --no-extract --no-gather-peak > gpurun_out/r02_c19_launches_target.log 2>&1
echo "launch list rc=$?"
B2="--steps 3 --no-compact --no-cpu-baseline --no-e2e --no-gather-peak --no-extract --npat 20000000"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_query_fused" --launch-skip 4 --launch-count 1 \
   -o /tmp/fused_defer -f python bench.py $B2 > gpurun_out/r02_c19_ncu_fused_defer.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_summary.py /tmp/fused_defer.ncu-rep gpurun_out/r02_c19_target_k_query_fused_defer_ncu.txt "k_query_fused_defer<FM,Q4,5,1> on the 1 GB DNA target (16-byte table entries), 20 M 32-mers, final round-2 sources (hash 15a4e14acfbb2704)"
python tools/ncu_hotspots.py /tmp/fused_defer.ncu-rep gpurun_out/r02_c19_target_k_query_fused_defer_hotspots.txt > /dev/null 2>&1
for wl in cfg3_rlfm cfg4_multi cfg5_bytes1g cfg2_dna100m cfg1_dna1m; do
  timeout 500 python bench.py --steps 10 --workload $wl > gpurun_out/r02_c19_bench_$wl.json 2> gpurun_out/r02_c19_bench_$wl.err
  echo "bench $wl rc=$?"; tail -c 300 gpurun_out/r02_c19_bench_$wl.err; head -c 260 gpurun_out/r02_c19_bench_$wl.json; echo
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c19_bench_reference_arm.json 2> gpurun_out/r02_c19_bench_reference_arm.err
echo "reference arm rc=$?"; head -c 300 gpurun_out/r02_c19_bench_reference_arm.json; echo
du -sh gpurun_out; ls -la gpurun_out/r02_c19*
echo done
