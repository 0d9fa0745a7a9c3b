#!/bin/bash
# Run each kernel-family check in its own process (a trapped kernel poisons the CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
for c in check_codebook_and_mask check_pool check_gate check_thin diag_conv check_conv_fwd check_conv_epilogues check_conv_wgrad; do
  timeout 600 python tools/gpu_bringup.py $c > gpurun_out/bringup_$c.log 2>&1
  echo "== $c exit $? =="
  grep -E "FAIL|EXCEPTION|Error|error|diag|FAILED rows|done in" gpurun_out/bringup_$c.log | head -40
done
