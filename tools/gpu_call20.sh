#!/bin/bash
# round-2 GPU call 20: k_offsets_emit (hit offsets + positions in one pass over the ranges): rich-path tests with it switched on,
# A/B on the target (one index, one batch, whole-batch equality, phases), default chosen ON THE BOX (same sed applied to the repo
# afterwards), then the captures of the final sources: ncu step capture -> ncu_traffic.json, default bench line, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
FMX_EMIT_FUSED=1 timeout 300 python -m pytest "tests/test_gpu_parity.py::test_rich_mode_query_parity[4-None-2]" "tests/test_gpu_parity.py::test_rich_mode_query_parity[4-None-0]" \
   tests/test_gpu_parity.py::test_rich_mode_many_hits_and_pipeline tests/test_gpu_parity.py::test_table_entries_with_text_context \
   tests/test_gpu_parity.py::test_packed_patterns tests/test_gpu_parity.py::test_scale_dna_420m_rich_and_compact \
   tests/test_gpu_parity.py::test_group_by_piece_and_replicated -q --timeout=300 --tb=short -rf -x > gpurun_out/r02_c20_pytest_subset.log 2>&1
T=$?
echo "pytest subset (FMX_EMIT_FUSED=1) rc=$T"; tail -6 gpurun_out/r02_c20_pytest_subset.log
timeout 300 python tools/ab_option.py --option emit_fused --values 0,1 > gpurun_out/r02_c20_ab_emit_fused.jsonl 2> gpurun_out/r02_c20_ab_emit_fused.err
echo "ab rc=$?"; cat gpurun_out/r02_c20_ab_emit_fused.jsonl; tail -c 300 gpurun_out/r02_c20_ab_emit_fused.err
ON=$(python - <<P
import json
rows = [json.loads(l) for l in open("gpurun_out/r02_c20_ab_emit_fused.jsonl") if l.startswith("{")]
s = [r for r in rows if "mean_ms" in r]
on = 0
if $T == 0 and s and s[0]["all_equal"]:
    m = s[0]["mean_ms"]
    if float(m["1"]) < 0.985 * float(m["0"]):
        on = 1
print(on)
P
)
echo "$ON" > gpurun_out/r02_c20_chosen.txt
echo "emit_fused default: $ON"
if [ "$ON" = "1" ]; then
  sed -i "s/^#define FMX_EMIT_FUSED_DEFAULT 0\$/#define FMX_EMIT_FUSED_DEFAULT 1/" fm-index_b200/csrc/phased.cuh
  grep -n "^#define FMX_EMIT_FUSED_DEFAULT" fm-index_b200/csrc/phased.cuh
  timeout 600 python -c "
import importlib.util
spec = importlib.util.spec_from_file_location('_b', 'fm-index_b200/build.py'); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b); print(b.build())" > gpurun_out/r02_c20_rebuild.log 2>&1
  echo "rebuild rc=$?"; tail -3 gpurun_out/r02_c20_rebuild.log
fi
python -c "import bench; print('source hash', bench.source_hash())"
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum"
timeout 300 ncu --metrics $M --clock-control none --profile-from-start off -o /tmp/step_rich -f python tools/prof_step.py --workload target_dna1g > gpurun_out/r02_c20_step_target_rich.log 2>&1
echo "ncu rich rc=$?"; grep "^{" gpurun_out/r02_c20_step_target_rich.log
python tools/ncu_traffic.py /tmp/step_rich.ncu-rep:gpurun_out/r02_c20_step_target_rich.log
cp profiles/ncu_traffic.json gpurun_out/r02_c20_ncu_traffic.json
ncu -i /tmp/step_rich.ncu-rep --page raw --csv > gpurun_out/r02_c20_step_target_rich_raw.csv 2>/dev/null
timeout 400 python bench.py > gpurun_out/r02_c20_bench_target_dna1g.json 2> gpurun_out/r02_c20_bench_target_dna1g.err
echo "bench default rc=$?"; tail -c 400 gpurun_out/r02_c20_bench_target_dna1g.err; head -c 300 gpurun_out/r02_c20_bench_target_dna1g.json; echo
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c20_launches_target.csv \
   python bench.py --steps 2 --warmup 3 --no-compact --no-cpu-baseline --no-e2e --no-extract --no-gather-peak > gpurun_out/r02_c20_launches_target.log 2>&1
echo "launch list rc=$?"
du -sh gpurun_out; ls -la gpurun_out/r02_c20*
echo done
