#!/bin/bash
# blocked 5-D tensor maps for the weight gradient (one TMA operation per operand): parity, then A/B against knob 17 = 0
mkdir -p gpurun_out
timeout 300 python tools/gpu_bringup.py check_conv_wgrad 2>&1 | tail -2
timeout 300 python tools/gpu_bringup.py check_conv_ex 2>&1 | tail -2
timeout 300 python tools/microbench.py wgrad > gpurun_out/r2_mb31_blk.log 2>&1
MCGM_DEBUG_SET=17=0 timeout 300 python tools/microbench.py wgrad > gpurun_out/r2_mb31_noblk.log 2>&1
paste -d"|" <(grep wgrad gpurun_out/r2_mb31_noblk.log | cut -c1-75) <(grep wgrad gpurun_out/r2_mb31_blk.log | cut -c38-75)
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_pixelcnn.py -q -x 2>&1 | tail -2
