#!/bin/bash
# round 2, call D1 (1 GPU): dK/dV micro-opt variants A/B, full single-GPU test suite, harness presets, small-N sweep, bench
mkdir -p gpurun_out
: > gpurun_out/r2d_ab.log
for lib in ab/*.so; do
  echo "== $lib" >> gpurun_out/r2d_ab.log
  FA_B200_LIB=$PWD/$lib timeout 200 python tests/bwd_ab.py 2>&1 | cut -c1-200 >> gpurun_out/r2d_ab.log
done
cat gpurun_out/r2d_ab.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log; tail -25 gpurun_out/r2d_pytest.log
(cd harness && timeout 300 ./flash_attn --config 3 > ../gpurun_out/r2d_harness_cfg3.log 2>&1; echo "rc=$?" >> ../gpurun_out/r2d_harness_cfg3.log)
cat gpurun_out/r2d_harness_cfg3.log
(cd harness && timeout 600 ./flash_attn --config 2 > ../gpurun_out/r2d_harness_cfg2.log 2>&1; echo "rc=$?" >> ../gpurun_out/r2d_harness_cfg2.log; cp benchmark_results_*.csv ../gpurun_out/ 2>/dev/null)
tail -60 gpurun_out/r2d_harness_cfg2.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench1.log 2> gpurun_out/r2d_bench1.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2d_bench1.log
