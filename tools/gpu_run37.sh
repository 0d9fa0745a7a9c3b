#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_pixelcnn.py -q -x 2>&1 | tail -2
timeout 200 python tools/microbench.py wgbias 2>&1 | grep want_bias
