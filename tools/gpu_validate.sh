#!/bin/bash
# validation pass on the final kernels (1 GPU): all GPU tests, smoke, every bench config, reference arm, step breakdown, ncu pass
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_t_final_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t_final_all.log
tail -4 gpurun_out/r2_t_final_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -1 gpurun_out/r2_smoke_final.log
for c in cifar10 omniglot mcvae mcglow; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2_bench_final_$c.log 2> gpurun_out/r2_bench_final_$c.err
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 --ref-seconds 30 > gpurun_out/r2_bench_final_ref.log 2> gpurun_out/r2_bench_final_ref.err
grep -h '"value"' gpurun_out/r2_bench_final_*.log | cut -c1-150
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown_final.log 2>&1
timeout 300 python tools/microbench.py conv wgrad gate thin epi > gpurun_out/r2_mb_final.log 2>&1
bash tools/ncu_pass_r02.sh > gpurun_out/r2_ncu_final.log 2>&1
tail -5 gpurun_out/r2_ncu_final.log
