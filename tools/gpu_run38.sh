#!/bin/bash
# last pass of the round: all GPU tests, the default bench line, step breakdown, refreshed weight-gradient captures
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t38_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t38_all.log; tail -3 gpurun_out/r2_t38_all.log
timeout 100 python bench.py --steps 20 --warmup 3 --no-torch-eager > gpurun_out/r2_bench38_cifar10.log 2> gpurun_out/r2_bench38_cifar10.err
grep -h '"value"' gpurun_out/r2_bench38_cifar10.log | cut -c1-150
timeout 60 python tools/step_breakdown.py > gpurun_out/r2_breakdown38.log 2>&1; head -3 gpurun_out/r2_breakdown38.log | cut -c1-120
NCU="ncu --clock-control none --kernel-name-base demangled"
P="python tools/profile_step.py"
F="--set full --import-source on --profile-from-start off"
timeout 60 $NCU $F -k regex:"conv_wgrad_kernel" -c 3 -o gpurun_out/prof_wgrad_d -f $P --mode d --warm 1 > gpurun_out/ncu_f3b.log 2>&1
timeout 60 $NCU $F -k regex:"conv_wgrad_kernel" -c 4 -o gpurun_out/prof_wgrad -f $P --mode g --warm 1 > gpurun_out/ncu_f3.log 2>&1
for r in gpurun_out/prof_wgrad*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > $b.raw.csv 2> /dev/null
  rm -f $r
done
