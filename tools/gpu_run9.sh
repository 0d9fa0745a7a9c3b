#!/bin/bash
# round-2 GPU pass 9: wgrad tap pairs as one N = 2*BN MMA — parity, microbench A/B, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q > gpurun_out/r2_k9.log 2>&1
echo "kernel tests rc=$?" >> gpurun_out/r2_k9.log
tail -3 gpurun_out/r2_k9.log
MCGM_DEBUG_SET="14=0" timeout 200 python tools/microbench.py wgrad > gpurun_out/r2_mb9_nopair.log 2>&1
timeout 200 python tools/microbench.py wgrad > gpurun_out/r2_mb9_pair.log 2>&1
paste -d'|' <(cut -c1-80 gpurun_out/r2_mb9_nopair.log) <(cut -c40-80 gpurun_out/r2_mb9_pair.log)
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_b9.log 2> gpurun_out/r2_b9.err
MCGM_DEBUG_SET="14=0" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager --no-kernel-profile > gpurun_out/r2_b9_nopair.log 2> gpurun_out/r2_b9_nopair.err
grep -h '"value"' gpurun_out/r2_b9*.log | cut -c1-170
