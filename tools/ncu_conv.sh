#!/bin/bash
# --set full captures of the big-shape conv launches (demangled-name regex selects the template instance)
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
timeout 300 python tools/profile_step.py --mode d > gpurun_out/plain_d0.log 2>&1 && \
$NCU --set full --import-source on -k regex:"conv_fwd_kernel<.int.256" -s 7 -c 2 -o gpurun_out/prof_conv_fwd -f python tools/profile_step.py --mode d > gpurun_out/ncu_full1.log 2>&1
$NCU --set full --import-source on -k regex:"conv_fwd_kernel<.int.128, .int.2" -s 0 -c 2 -o gpurun_out/prof_conv_fwd_d -f python tools/profile_step.py --mode d > gpurun_out/ncu_full1b.log 2>&1
timeout 300 python tools/profile_step.py --mode g > gpurun_out/plain_g0.log 2>&1 && \
$NCU --set full --import-source on -k regex:"conv_wgrad_kernel<.int.256" -s 1 -c 2 -o gpurun_out/prof_wgrad -f python tools/profile_step.py --mode g > gpurun_out/ncu_full3.log 2>&1
for f in gpurun_out/ncu_full1.log gpurun_out/ncu_full1b.log gpurun_out/ncu_full3.log; do grep -E "No kernels|Report" $f | head -2; done
