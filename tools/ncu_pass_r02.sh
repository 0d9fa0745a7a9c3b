#!/bin/bash
# ncu evidence, round 2: launch lists of one D update + one G update (steady state) and --set full captures of the
# hot kernels (B200_PROFILING.md recipe: --clock-control none, each capture only after the same command exited 0 without ncu)
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
P="python tools/profile_step.py"
timeout 300 $P --mode d --warm 1 > gpurun_out/plain_d.log 2>&1 && \
$NCU --profile-from-start off --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_d.csv $P --mode d --warm 1 > gpurun_out/ncu_d.log 2>&1
timeout 300 $P --mode g --warm 1 > gpurun_out/plain_g.log 2>&1 && \
$NCU --profile-from-start off --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_g.csv $P --mode g --warm 1 > gpurun_out/ncu_g.log 2>&1
F="--set full --import-source on --profile-from-start off"
$NCU $F -k regex:"conv_fwd_kernel<.int.256" -c 4 -o gpurun_out/prof_conv_fwd -f $P --mode g --warm 1 > gpurun_out/ncu_f1.log 2>&1
$NCU --set full --profile-from-start off -k regex:"conv_fwd_kernel<.int.256" -s 7 -c 3 -o gpurun_out/prof_conv_fwd_big -f $P --mode g --warm 1 > gpurun_out/ncu_f1b.log 2>&1
$NCU $F -k regex:"conv_fwd_kernel<.int.128, .int.2" -c 3 -o gpurun_out/prof_conv_fwd_d -f $P --mode d --warm 1 > gpurun_out/ncu_f2.log 2>&1
$NCU $F -k regex:"conv_wgrad_kernel" -c 4 -o gpurun_out/prof_wgrad -f $P --mode g --warm 1 > gpurun_out/ncu_f3.log 2>&1
$NCU $F -k regex:"conv_wgrad_kernel" -c 3 -o gpurun_out/prof_wgrad_d -f $P --mode d --warm 1 > gpurun_out/ncu_f3b.log 2>&1
$NCU $F -k regex:"gate_fwd_kernel|gate_bwd_reduce_kernel|gate_bwd_apply_kernel" -c 8 -o gpurun_out/prof_gate -f $P --mode g --warm 1 > gpurun_out/ncu_f4.log 2>&1
$NCU $F -k regex:"thin_conv_fwd_umma_kernel|thin_conv_wgrad_umma_kernel" -c 4 -o gpurun_out/prof_thin -f $P --mode d --warm 1 > gpurun_out/ncu_f5.log 2>&1
for f in gpurun_out/ncu_f*.log; do echo $f; grep -E "No kernels|Report|==ERROR" $f | head -2; done
# gpurun brings back at most 64 MiB: export what the summaries need as CSV on the box and drop the reports
for r in gpurun_out/prof_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > $b.raw.csv 2> /dev/null
  ncu -i $r --page source --csv --print-source sass > $b.source.csv 2> /dev/null
  # keep the source page small: the 400 hottest lines by sampled stalls would need parsing; keep it whole but gzip
  gzip -f $b.source.csv
done
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/prof_*.ncu-rep
du -sh gpurun_out
