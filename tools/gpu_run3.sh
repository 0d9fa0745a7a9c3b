#!/bin/bash
# round-2 GPU pass 3: cluster-multicast forward conv — parity, microbench A/B, full tests, bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q > gpurun_out/r2_k3.log 2>&1
echo "kernel tests rc=$?" >> gpurun_out/r2_k3.log
tail -4 gpurun_out/r2_k3.log
for cl in 1 2 4; do
  MCGM_DEBUG_SET="11=$cl" timeout 300 python tools/microbench.py conv epi > gpurun_out/r2_mb3_cl$cl.log 2>&1
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t3_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t3_all.log
tail -8 gpurun_out/r2_t3_all.log
for cl in 1 2 4; do
  MCGM_DEBUG_SET="11=$cl" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_b3_cl$cl.log 2> gpurun_out/r2_b3_cl$cl.err
done
MCGM_BN_STATS_IN_CONV=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_b3_nostats.log 2> gpurun_out/r2_b3_nostats.err
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown3.log 2>&1
grep -h '"value"' gpurun_out/r2_b3_*.log | cut -c1-160
