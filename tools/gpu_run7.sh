#!/bin/bash
# round-2 GPU pass 7: validate (one statistics slot per CTA, thin kernels two per SM), bench, then the ncu evidence
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_b7.log 2> gpurun_out/r2_b7.err
grep -h '"value"' gpurun_out/r2_b7.log | cut -c1-170
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown7.log 2>&1
bash tools/ncu_pass_r02.sh > gpurun_out/r2_ncu_pass.log 2>&1
tail -12 gpurun_out/r2_ncu_pass.log
