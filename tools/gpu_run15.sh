#!/bin/bash
# A/B: double-buffered epilogue staging (128-wide pair tiles; optionally the 256-wide tiles), side slab by TMA or through registers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/r2_t15_kernels.log 2>&1; tail -3 gpurun_out/r2_t15_kernels.log
run() { # name lib debug
  MCGM_LIB_PATH=$2 MCGM_DEBUG_SET=$3 timeout 300 python tools/microbench.py conv epi > gpurun_out/r2_mb15_$1.log 2>&1
  MCGM_LIB_PATH=$2 MCGM_DEBUG_SET=$3 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_bench15_$1.log 2>/dev/null
  echo "== $1"; grep -h '"value"' gpurun_out/r2_bench15_$1.log | cut -c1-130
}
run nb1 $PWD/abvariants/libmcgm_nb1.so ""
run nb2 "" ""
run nb2_sidereg "" "10=3"
run nb2_256 $PWD/abvariants/libmcgm_nb2_256.so ""
run nb2_256_sidereg $PWD/abvariants/libmcgm_nb2_256.so "10=3"
paste <(cut -c1-52 gpurun_out/r2_mb15_nb1.log) <(cut -c44-52 gpurun_out/r2_mb15_nb2.log) <(cut -c44-52 gpurun_out/r2_mb15_nb2_sidereg.log) <(cut -c44-52 gpurun_out/r2_mb15_nb2_256.log) <(cut -c44-52 gpurun_out/r2_mb15_nb2_256_sidereg.log)
