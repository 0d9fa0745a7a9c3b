#!/bin/bash
# quick diagnostic capture (source-level samples) of the discriminator's conv and weight-gradient kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
P="python tools/profile_step.py"
F="--set full --import-source on --profile-from-start off"
timeout 300 $P --mode d --warm 1 > gpurun_out/plain_d.log 2>&1
$NCU $F -k regex:"conv_fwd_kernel<.int.128, .int.2" -c 2 -o gpurun_out/q_conv_fwd_d -f $P --mode d --warm 1 > gpurun_out/ncu_q1.log 2>&1
$NCU $F -k regex:"conv_wgrad_kernel" -c 2 -o gpurun_out/q_wgrad_d -f $P --mode d --warm 1 > gpurun_out/ncu_q2.log 2>&1
for r in gpurun_out/q_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > $b.raw.csv 2> /dev/null
  ncu -i $r --page source --csv --print-source sass > $b.source.csv 2> /dev/null
  gzip -f $b.source.csv
done
rm -f gpurun_out/q_*.ncu-rep
