#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -k "two_rank" > gpurun_out/r2_t10_ddp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t10_ddp.log
grep -E "launch form|rank [01]:|2-rank NCCL|passed|failed|^E  " gpurun_out/r2_t10_ddp.log | cut -c1-400
