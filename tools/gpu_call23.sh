#!/bin/bash
# round-2 GPU call 23: per-step ncu capture of the COMPACT target index on the final sources (completes ncu_traffic.json)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum"
timeout 100 ncu --metrics $M --clock-control none --profile-from-start off -o /tmp/step_compact -f python tools/prof_step.py --workload target_dna1g --mode compact > gpurun_out/r02_c23_step_target_compact.log 2>&1
echo "ncu compact rc=$?"; grep "^{" gpurun_out/r02_c23_step_target_compact.log
python tools/ncu_traffic.py /tmp/step_compact.ncu-rep:gpurun_out/r02_c23_step_target_compact.log
cp profiles/ncu_traffic.json gpurun_out/r02_c23_ncu_traffic.json
