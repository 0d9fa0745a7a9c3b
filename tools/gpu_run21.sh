#!/bin/bash
# CTA-pair (cta_group::2) forward conv: parity, then A/B against single-CTA MMAs (knob 16)
mkdir -p gpurun_out
timeout 300 python tools/gpu_bringup.py check_conv_fwd > gpurun_out/r2_t21_fwd.log 2>&1; echo "check_conv_fwd rc=$?"; tail -5 gpurun_out/r2_t21_fwd.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/r2_t21_kernels.log 2>&1; tail -3 gpurun_out/r2_t21_kernels.log
timeout 300 python tools/microbench.py conv epi > gpurun_out/r2_mb21_pair.log 2>&1
MCGM_DEBUG_SET=16=0 timeout 300 python tools/microbench.py conv epi > gpurun_out/r2_mb21_single.log 2>&1
paste -d'|' <(grep k3 gpurun_out/r2_mb21_single.log | cut -c1-72) <(grep k3 gpurun_out/r2_mb21_pair.log | cut -c38-72)
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_bench21.log 2>/dev/null
grep -h '"value"' gpurun_out/r2_bench21.log | cut -c1-130
