#!/bin/bash
# elect.sync issue loops: kernel parity, microbenchmarks, bench, step breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/r2_t13_kernels.log 2>&1; tail -3 gpurun_out/r2_t13_kernels.log
timeout 300 python tools/microbench.py thin > gpurun_out/r2_mb13_thin.log 2>&1; cat gpurun_out/r2_mb13_thin.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_bench13.log 2> gpurun_out/r2_bench13.err
grep -h '"value"' gpurun_out/r2_bench13.log | cut -c1-200
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown13.log 2>&1; head -30 gpurun_out/r2_breakdown13.log | cut -c1-150
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t13_all.log 2>&1; tail -3 gpurun_out/r2_t13_all.log
