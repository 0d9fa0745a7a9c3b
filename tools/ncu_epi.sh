#!/bin/bash
# where does the staged epilogue spend its time?  epilogue-only launches (debug knob 15 = 7, single-CTA MMAs) under ncu
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled --set full --import-source on"
for fl in 0x0 0x3; do
  MCGM_DEBUG_SET="15=7,16=0" timeout 200 python tools/one_conv.py 32 128 $fl > gpurun_out/one_conv.log 2>&1
  MCGM_DEBUG_SET="15=7,16=0" $NCU -k regex:"conv_fwd_kernel" -s 2 -c 1 -o gpurun_out/q_epi_$fl -f python tools/one_conv.py 32 128 $fl > gpurun_out/ncu_qe.log 2>&1
  ncu -i gpurun_out/q_epi_$fl.ncu-rep --page raw --csv > gpurun_out/q_epi_$fl.raw.csv 2> /dev/null
  ncu -i gpurun_out/q_epi_$fl.ncu-rep --page source --csv --print-source sass > gpurun_out/q_epi_$fl.source.csv 2> /dev/null
  gzip -f gpurun_out/q_epi_$fl.source.csv
done
rm -f gpurun_out/q_epi_*.ncu-rep
tail -3 gpurun_out/ncu_qe.log
