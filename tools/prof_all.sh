set -x
# tools/gather_bench > gpurun_out/gather_bench.jsonl   (random-access study, see profiles/r01b_random_access_study.md)
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
$NCU -k regex:'k_search|k_locate' -f -o gpurun_out/prof_target_r1b python tools/prof_step.py --workload target_dna1g --npat 20000000 > gpurun_out/prof_target_r1b.log 2>&1
$NCU -k regex:'k_search|k_locate' -f -o gpurun_out/prof_cfg5_r1b python tools/prof_step.py --workload cfg5_bytes1g --npat 10000000 > gpurun_out/prof_cfg5_r1b.log 2>&1
$NCU -k regex:'k_search|k_locate' -f -o gpurun_out/prof_cfg3_r1b python tools/prof_step.py --workload cfg3_rlfm --npat 2000000 > gpurun_out/prof_cfg3_r1b.log 2>&1
$NCU -k regex:'k_search|k_locate' -f -o gpurun_out/prof_cfg2_r1b python tools/prof_step.py --workload cfg2_dna100m > gpurun_out/prof_cfg2_r1b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg2_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gather-peak > gpurun_out/launches_cfg2_r1b.log 2>&1
ls -la gpurun_out/*.ncu-rep
