#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 300 python tools/profile_step.py --mode d --warm 1 > gpurun_out/plain_d.log 2>&1 && \
$NCU --profile-from-start off --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_d.csv python tools/profile_step.py --mode d --warm 1 > gpurun_out/ncu_d.log 2>&1
timeout 300 python tools/profile_step.py --mode g --warm 1 > gpurun_out/plain_g.log 2>&1 && \
$NCU --profile-from-start off --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_g.csv python tools/profile_step.py --mode g --warm 1 > gpurun_out/ncu_g.log 2>&1
