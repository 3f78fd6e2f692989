#!/bin/bash
# the driver's SCALE invocation at N=2 (largest by-piece partitions beside the resident target index)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_final_bench_2gpu.json 2> gpurun_out/r02_final_bench_2gpu.err
echo "bench 2gpu rc=$?"; tail -c 600 gpurun_out/r02_final_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference_arm_2gpu.json 2> gpurun_out/r02_final_bench_reference_arm_2gpu.err
echo "reference arm 2gpu rc=$?"; tail -c 300 gpurun_out/r02_final_bench_reference_arm_2gpu.json
