#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gpu_bringup.py check_conv_wgrad 2>&1 | tail -2
MCGM_DEBUG_SET=18=0 timeout 200 python tools/microbench.py wgbias 2>&1 | grep want_bias | sed 's/^/group0: /'
timeout 200 python tools/microbench.py wgbias 2>&1 | grep want_bias | sed 's/^/last:   /'
