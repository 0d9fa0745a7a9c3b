#!/bin/bash
# round-2 GPU pass 6: thin (RGB-side) convs on tcgen05 / TMA — parity, microbench A/B, full tests, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q > gpurun_out/r2_k6.log 2>&1
echo "kernel tests rc=$?" >> gpurun_out/r2_k6.log
tail -4 gpurun_out/r2_k6.log
MCGM_DEBUG_SET="12=0,13=0" timeout 200 python tools/microbench.py thin > gpurun_out/r2_mb6_old.log 2>&1
timeout 200 python tools/microbench.py thin > gpurun_out/r2_mb6_new.log 2>&1
cat gpurun_out/r2_mb6_old.log gpurun_out/r2_mb6_new.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_t6_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t6_all.log
tail -6 gpurun_out/r2_t6_all.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_b6.log 2> gpurun_out/r2_b6.err
MCGM_DEBUG_SET="12=0,13=0" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager --no-kernel-profile > gpurun_out/r2_b6_oldthin.log 2> gpurun_out/r2_b6_oldthin.err
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown6.log 2>&1
grep -h '"value"' gpurun_out/r2_b6*.log | cut -c1-170
