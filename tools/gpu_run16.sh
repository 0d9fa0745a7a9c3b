#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/r2_t16_kernels.log 2>&1; tail -3 gpurun_out/r2_t16_kernels.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_bench16.log 2>/dev/null
grep -h '"value"' gpurun_out/r2_bench16.log | cut -c1-130
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown16.log 2>&1; grep -E "eager|col2im|bn_finalize" gpurun_out/r2_breakdown16.log | cut -c1-150
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -q -x > gpurun_out/r2_t16_parity.log 2>&1; tail -3 gpurun_out/r2_t16_parity.log
