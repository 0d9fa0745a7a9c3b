#!/bin/bash
# A/B: 32-element k-blocks (64-byte rows, SWIZZLE_64B, 8 ring stages) for the 128-wide pair tile / the 256-wide tile
mkdir -p gpurun_out
for v in kb32_128 kb32_both; do
  MCGM_LIB_PATH=$PWD/abvariants/libmcgm_$v.so timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/r2_t20_$v.log 2>&1; echo "== $v"; tail -2 gpurun_out/r2_t20_$v.log
  MCGM_LIB_PATH=$PWD/abvariants/libmcgm_$v.so timeout 300 python tools/microbench.py conv epi legs > gpurun_out/r2_mb20_$v.log 2>&1
  MCGM_LIB_PATH=$PWD/abvariants/libmcgm_$v.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_bench20_$v.log 2>/dev/null
  grep -h '"value"' gpurun_out/r2_bench20_$v.log | cut -c1-130
done
timeout 300 python tools/microbench.py conv epi > gpurun_out/r2_mb20_base.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2_bench20_base.log 2>/dev/null
grep -h '"value"' gpurun_out/r2_bench20_base.log | cut -c1-130
for v in base kb32_128 kb32_both; do echo "== $v"; grep -E "k3" gpurun_out/r2_mb20_$v.log | grep -v legs | cut -c1-90; done
grep legs gpurun_out/r2_mb20_kb32_both.log
