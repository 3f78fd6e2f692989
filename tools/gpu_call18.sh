#!/bin/bash
# round-2 GPU call 18: A/B of the fused query kernel's occupancy / queue variants on the target (one index, one batch,
# whole-batch equality), the winner becomes the default instance ON THE BOX (same sed is applied to the repo afterwards,
# so the source hash of the captures matches), then: ncu step capture, default bench line, launch list, ncu --set full,
# config 3 with the new AUTO policy
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout=300 --tb=short -rf -x -k "table_entries or packed or rich_mode_many" > gpurun_out/r02_c18_pytest_subset.log 2>&1
echo "pytest subset rc=$?"; tail -5 gpurun_out/r02_c18_pytest_subset.log
timeout 400 python tools/ab_defer.py > gpurun_out/r02_c18_ab_defer.jsonl 2> gpurun_out/r02_c18_ab_defer.err
echo "ab rc=$?"; cat gpurun_out/r02_c18_ab_defer.jsonl; tail -c 300 gpurun_out/r02_c18_ab_defer.err
read BEST B R G < <(python - <<'P'
import json
rows = [json.loads(l) for l in open("gpurun_out/r02_c18_ab_defer.jsonl") if l.startswith("{")]
rows = [r for r in rows if "fused_defer" in r]
base = [r for r in rows if r["fused_defer"] == 4 and r["blocks_per_sm"] == 64]
best = (4, 64)
if base and all(r["same_offsets_and_positions_as_first_variant"] in (None, True) for r in rows):
    cand = min((r for r in rows if r["fused_defer"] in (3, 4, 5, 6, 7)), key=lambda r: r["ms_per_step"])
    if cand["ms_per_step"] < 0.98 * base[0]["ms_per_step"]:
        best = (cand["fused_defer"], cand["blocks_per_sm"])
b, r = {3: (6, 1), 4: (5, 1), 5: (5, 2), 6: (6, 2), 7: (5, 4)}[best[0]]
print(best[0], b, r, best[1])
P
)
echo "$BEST $B $R $G" > gpurun_out/r02_c18_chosen.txt
echo "chosen variant $BEST: blocks $B rounds $R grid blocks per SM $G"
if [ "$B$R$G" != "5164" ]; then
  sed -i "s/^#define FMX_DEFER_BLOCKS 5\$/#define FMX_DEFER_BLOCKS $B/; s/^#define FMX_DEFER_ROUNDS 1\$/#define FMX_DEFER_ROUNDS $R/; s/^#define FMX_QUERY_BLOCKS_PER_SM 64\$/#define FMX_QUERY_BLOCKS_PER_SM $G/" fm-index_b200/csrc/phased.cuh
  grep -n "^#define FMX_DEFER_" fm-index_b200/csrc/phased.cuh
  timeout 600 python -c "
import importlib.util
spec = importlib.util.spec_from_file_location('_b', 'fm-index_b200/build.py'); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b); print(b.build())" > gpurun_out/r02_c18_rebuild.log 2>&1
  echo "rebuild rc=$?"; tail -3 gpurun_out/r02_c18_rebuild.log
  timeout 700 python -m pytest tests -m gpu -q --timeout=600 --tb=short -rf > gpurun_out/r02_c18_pytest_full.log 2>&1
  echo "pytest full (rebuilt) rc=$?"; tail -5 gpurun_out/r02_c18_pytest_full.log
fi
python -c "import bench; print('source hash', bench.source_hash())"
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum"
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -o gpurun_out/r02_c18_step_target_rich -f python tools/prof_step.py --workload target_dna1g > gpurun_out/r02_c18_step_target_rich.log 2>&1
echo "ncu rich rc=$?"; grep "^{" gpurun_out/r02_c18_step_target_rich.log
python tools/ncu_traffic.py gpurun_out/r02_c18_step_target_rich.ncu-rep:gpurun_out/r02_c18_step_target_rich.log   # the box's copy: the line below carries the capture
timeout 600 python bench.py > gpurun_out/r02_c18_bench_target_dna1g.json 2> gpurun_out/r02_c18_bench_target_dna1g.err
echo "bench default rc=$?"; tail -c 400 gpurun_out/r02_c18_bench_target_dna1g.err; head -c 300 gpurun_out/r02_c18_bench_target_dna1g.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c18_launches_target.csv \
   python bench.py --steps 2 --warmup 3 --no-compact --no-cpu-baseline --no-e2e --no-extract --no-gather-peak > gpurun_out/r02_c18_launches_target.log 2>&1
echo "launch list rc=$?"
B2="--steps 3 --no-compact --no-cpu-baseline --no-e2e --no-gather-peak --no-extract --npat 20000000"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_query_fused" --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r02_c18_fused_defer -f python bench.py $B2 > gpurun_out/r02_c18_ncu_fused_defer.log 2>&1
echo "ncu full rc=$?"
timeout 500 python bench.py --steps 10 --workload cfg3_rlfm > gpurun_out/r02_c18_bench_cfg3_rlfm.json 2> gpurun_out/r02_c18_bench_cfg3_rlfm.err
echo "bench cfg3 rc=$?"; tail -c 300 gpurun_out/r02_c18_bench_cfg3_rlfm.err; head -c 300 gpurun_out/r02_c18_bench_cfg3_rlfm.json; echo
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -o gpurun_out/r02_c18_step_target_compact -f python tools/prof_step.py --workload target_dna1g --mode compact > gpurun_out/r02_c18_step_target_compact.log 2>&1
echo "ncu compact rc=$?"
ls -la gpurun_out/r02_c18*
echo done
