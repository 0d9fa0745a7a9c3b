#!/bin/bash
# round-2 GPU pass 1: parity tests, every bench config, per-shape breakdown
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_env.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_ddp.py::test_two_rank_nccl_iteration_vs_oracle_shards > gpurun_out/r2_t1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t1.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t1_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t1_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_cifar10.log 2> gpurun_out/r2_bench_cifar10.err
for c in omniglot mcvae mcglow; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_$c.log 2> gpurun_out/r2_bench_$c.err
done
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown1.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
tail -5 gpurun_out/r2_t1_all.log
cat gpurun_out/r2_bench_cifar10.log | cut -c1-600
