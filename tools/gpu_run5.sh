#!/bin/bash
# round-2 GPU pass 5 (2 GPUs): NCCL parity of the data-parallel forms + N=2 bench of both launch forms
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_env2.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_ddp.py -q -x > gpurun_out/r2_t5_ddp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t5_ddp.log
tail -6 gpurun_out/r2_t5_ddp.log
for mode in ingraph segmented; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --ddp $mode > gpurun_out/r2_b5_n2_$mode.log 2> gpurun_out/r2_b5_n2_$mode.err
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_b5_n1.log 2> gpurun_out/r2_b5_n1.err
grep -h '"value"' gpurun_out/r2_b5_*.log | cut -c1-170
tail -3 gpurun_out/r2_b5_n2_ingraph.err
