#!/bin/bash
# ncu evidence for the round: launch lists of one D update + one G update (steady state: after one
# warm-up update), and --set full captures of the big-shape launches of each hot kernel.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 300 python tools/profile_step.py --mode d --warm 1 > gpurun_out/plain_d.log 2>&1 && \
$NCU --profile-from-start off --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_d.csv python tools/profile_step.py --mode d --warm 1 > gpurun_out/ncu_d.log 2>&1
timeout 300 python tools/profile_step.py --mode g --warm 1 > gpurun_out/plain_g.log 2>&1 && \
$NCU --profile-from-start off --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_g.csv python tools/profile_step.py --mode g --warm 1 > gpurun_out/ncu_g.log 2>&1
timeout 300 python tools/profile_step.py --mode d > gpurun_out/plain_d0.log 2>&1


$NCU --set full -k regex:"gate_fwd_kernel|bn_partial_kernel" -s 12 -c 4 -o gpurun_out/prof_gate_fwd python tools/profile_step.py --mode d > gpurun_out/ncu_full2.log 2>&1
timeout 300 python tools/profile_step.py --mode g > gpurun_out/plain_g0.log 2>&1 && \
bash tools/ncu_conv.sh
$NCU --set full -k regex:"gate_bwd_reduce_kernel|gate_bwd_apply_kernel" -s 4 -c 2 -o gpurun_out/prof_gate_bwd python tools/profile_step.py --mode g > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out | head -30
