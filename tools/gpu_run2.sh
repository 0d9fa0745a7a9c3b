#!/bin/bash
# round-2 GPU pass 2: staged TMA-store epilogue — kernel parity under each knob setting, microbench A/B, debug scripts
mkdir -p gpurun_out
K="tests/test_gpu_kernels.py"
for knob in 7 3 1 0; do
  MCGM_DEBUG_SET="10=$knob" timeout 600 python -m pytest $K -q -x -k "conv or thin or batched" > gpurun_out/r2_k_$knob.log 2>&1
  echo "knob $knob rc=$?" >> gpurun_out/r2_k_$knob.log
  tail -3 gpurun_out/r2_k_$knob.log
done
for knob in 0 1 7; do
  MCGM_DEBUG_SET="10=$knob" timeout 300 python tools/microbench.py conv epi > gpurun_out/r2_mb_$knob.log 2>&1
done
timeout 300 python tools/debug_segmented.py > gpurun_out/r2_dbg_seg.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t2_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t2_all.log
tail -8 gpurun_out/r2_t2_all.log
for knob in 0 7; do
  MCGM_DEBUG_SET="10=$knob" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_b2_knob$knob.log 2> gpurun_out/r2_b2_knob$knob.err
done
MCGM_BN_STATS_IN_CONV=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_b2_stats.log 2> gpurun_out/r2_b2_stats.err
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown2.log 2>&1
timeout 600 python bench.py --config omniglot --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager --no-kernel-profile > gpurun_out/r2_b2_omni.log 2> gpurun_out/r2_b2_omni.err
grep -h '"value"' gpurun_out/r2_b2_*.log | cut -c1-200
