#!/bin/bash
# 2 GPUs: DDP parity tests + the bench as the driver launches it (torchrun), also the reference arm's N>1 contract
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddp.py -q > gpurun_out/r2_t28_ddp.log 2>&1; tail -2 gpurun_out/r2_t28_ddp.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b28_n2.log 2> gpurun_out/r2_b28_n2.err
echo "bench rc=$?"; grep -h '"value"' gpurun_out/r2_b28_n2.log | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --ref-seconds 10 > gpurun_out/r2_b28_ref2.log 2> gpurun_out/r2_b28_ref2.err
echo "ref rc=$?"; grep -h '"impl"' gpurun_out/r2_b28_ref2.log | cut -c1-200
