#!/bin/bash
# round-2 GPU call 24: per-step ncu captures of configs 4 and 5 on the final sources (ncu_traffic.json entries for their bench lines)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum"
for wl in cfg4_multi cfg5_bytes1g; do
  timeout 45 ncu --metrics $M --clock-control none --profile-from-start off -o /tmp/step_$wl -f python tools/prof_step.py --workload $wl > gpurun_out/r02_c24_step_$wl.log 2>&1
  echo "ncu $wl rc=$?"; grep "^{" gpurun_out/r02_c24_step_$wl.log
  [ -f /tmp/step_$wl.ncu-rep ] && python tools/ncu_traffic.py /tmp/step_$wl.ncu-rep:gpurun_out/r02_c24_step_$wl.log > /dev/null
  cp profiles/ncu_traffic.json gpurun_out/r02_c24_ncu_traffic.json
done
python -c "
import json
for k,v in json.load(open('gpurun_out/r02_c24_ncu_traffic.json')).items(): print(k, v['source_hash'], v['npat'], round(v['dram_bytes_per_step']/1e9,2), round(v['l2_read_requests_per_step']/1e6,1), v['active_lanes_per_instruction'])"
