#!/bin/bash
# round-2 final validation pass (1 GPU): all GPU tests, smoke, thin microbench, every bench config, reference arm, sanitizer, one more ncu capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_t12_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t12_all.log
tail -5 gpurun_out/r2_t12_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke12.log 2>&1; tail -2 gpurun_out/r2_smoke12.log
timeout 200 python tools/microbench.py thin > gpurun_out/r2_mb12_thin.log 2>&1; cat gpurun_out/r2_mb12_thin.log
for c in cifar10 omniglot mcvae mcglow; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2_bench12_$c.log 2> gpurun_out/r2_bench12_$c.err
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 --ref-seconds 30 > gpurun_out/r2_bench12_ref.log 2> gpurun_out/r2_bench12_ref.err
grep -h '"value"' gpurun_out/r2_bench12_*.log | cut -c1-150
timeout 300 python tools/step_breakdown.py > gpurun_out/r2_breakdown12.log 2>&1
NCU="ncu --clock-control none --kernel-name-base demangled"
timeout 300 python tools/profile_step.py --mode g --warm 1 > gpurun_out/plain_g2.log 2>&1 && \
$NCU --set full --profile-from-start off -k regex:"conv_fwd_kernel<.int.256" -s 7 -c 3 -o gpurun_out/prof_conv_fwd_big -f python tools/profile_step.py --mode g --warm 1 > gpurun_out/ncu_f6.log 2>&1
ncu -i gpurun_out/prof_conv_fwd_big.ncu-rep --page raw --csv > gpurun_out/prof_conv_fwd_big.raw.csv 2>/dev/null; rm -f gpurun_out/prof_conv_fwd_big.ncu-rep
bash tools/sanitize.sh > gpurun_out/r2_sanitize.log 2>&1
tail -14 gpurun_out/r2_sanitize.log
du -sh gpurun_out
