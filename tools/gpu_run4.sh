#!/bin/bash
# A/B of library build variants on one box: microbench conv/epi, then the bench with the candidates
mkdir -p gpurun_out
R=$PWD
for v in main keepold r200 ring224; do
  if [ $v = main ]; then L=""; else L="$R/abvariants/libmcgm_$v.so"; fi
  MCGM_LIB_PATH=$L MCGM_DEBUG_SET="11=1" timeout 300 python tools/microbench.py conv epi > gpurun_out/r2_mb4_$v.log 2>&1
done
for v in main keepold ring224; do
  if [ $v = main ]; then L=""; else L="$R/abvariants/libmcgm_$v.so"; fi
  MCGM_LIB_PATH=$L MCGM_DEBUG_SET="11=1" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-eager --no-kernel-profile > gpurun_out/r2_b4_$v.log 2> gpurun_out/r2_b4_$v.err
done
grep -h '"value"' gpurun_out/r2_b4_*.log | cut -c1-160
