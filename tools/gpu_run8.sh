#!/bin/bash
# round-2 GPU pass 8 (2 GPUs): NCCL parity of the data-parallel forms (teardown fixed) + N=2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddp.py -q -k "two_rank" > gpurun_out/r2_t8_ddp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t8_ddp.log
grep -E "launch form|rank [01]:|2-rank NCCL|passed|failed" gpurun_out/r2_t8_ddp.log | cut -c1-220
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b8_n2.log 2> gpurun_out/r2_b8_n2.err
echo "bench rc=$?"
grep -h '"value"' gpurun_out/r2_b8_n2.log | cut -c1-170
tail -3 gpurun_out/r2_b8_n2.err
