#!/bin/bash
# round-2 GPU final 8-GPU call: the driver's SCALE invocation at N=8 -- target workload replicated + config 4 piece-partitioned
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_final_bench_8gpu.json 2> gpurun_out/r02_final_bench_8gpu.err
echo "bench 8gpu rc=$?"; tail -c 600 gpurun_out/r02_final_bench_8gpu.err; head -c 1500 gpurun_out/r02_final_bench_8gpu.json
