#!/bin/bash
# 8-GPU check of the in-graph NCCL step (what the driver's scaling run launches)
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 --no-kernel-profile > gpurun_out/r2_b34_n8.log 2> gpurun_out/r2_b34_n8.err
echo "bench rc=$?"
grep -h '"value"' gpurun_out/r2_b34_n8.log | cut -c1-400
tail -4 gpurun_out/r2_b34_n8.err | cut -c1-300
