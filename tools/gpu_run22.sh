#!/bin/bash
# CTA pairs + double-buffered staging (128-wide by default; 256-wide as a variant): parity, A/B, legs, step breakdown
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/r2_t22_kernels.log 2>&1; tail -3 gpurun_out/r2_t22_kernels.log
timeout 300 python tools/microbench.py conv epi legs > gpurun_out/r2_mb22_nb2.log 2>&1
MCGM_LIB_PATH=$PWD/abvariants/libmcgm_p256nb2.so timeout 300 python tools/microbench.py conv epi > gpurun_out/r2_mb22_p256nb2.log 2>&1
paste -d'|' <(grep k3 gpurun_out/r2_mb22_nb2.log | grep -v legs | cut -c1-72) <(grep k3 gpurun_out/r2_mb22_p256nb2.log | cut -c38-72)
grep legs gpurun_out/r2_mb22_nb2.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_bench22.log 2>/dev/null
grep -h '"value"' gpurun_out/r2_bench22.log | cut -c1-130
MCGM_LIB_PATH=$PWD/abvariants/libmcgm_p256nb2.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_bench22b.log 2>/dev/null
grep -h '"value"' gpurun_out/r2_bench22b.log | cut -c1-130
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown22.log 2>&1; head -24 gpurun_out/r2_breakdown22.log | cut -c1-150
