#!/bin/bash
# 2 GPUs: DDP parity with the pipelined in-graph step, then the bench with / without the overlap (MCGM_DDP_PIPELINE)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddp.py -q > gpurun_out/r2_t33_ddp.log 2>&1; tail -2 gpurun_out/r2_t33_ddp.log
for v in 1 0; do
  MCGM_DDP_PIPELINE=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$v bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager --no-kernel-profile > gpurun_out/r2_b33_n2_p$v.log 2> gpurun_out/r2_b33_n2_p$v.err
  echo "pipeline=$v rc=$?"; grep -h '"value"' gpurun_out/r2_b33_n2_p$v.log | cut -c1-140
done
