#!/bin/bash
# round 2, call C (2 GPUs): fused backward variants A/B + wait trace; mgpu stall diagnosis
mkdir -p gpurun_out
: > gpurun_out/r2c_ab.log
for lib in ab/*.so; do
  echo "== $lib" >> gpurun_out/r2c_ab.log
  FA_B200_LIB=$PWD/$lib timeout 200 python tests/bwd_ab.py >> gpurun_out/r2c_ab.log 2>&1
done
cat gpurun_out/r2c_ab.log
for d in 128 64; do FA_B200_LIB=$PWD/ab_trace.so timeout 120 python tests/fused_probe.py $d 1; done > gpurun_out/r2c_trace.log 2>&1
FA_B200_LIB=$PWD/ab_trace.so timeout 120 python tests/fused_probe.py 128 0 >> gpurun_out/r2c_trace.log 2>&1
cat gpurun_out/r2c_trace.log
for cfg in "-1 0" "-1 1" "2 1" "1 1"; do
  echo "== mgpu_debug $cfg" >> gpurun_out/r2c_mgpu.log
  timeout 60 python tests/mgpu_debug.py $cfg >> gpurun_out/r2c_mgpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_mgpu.log
done
cat gpurun_out/r2c_mgpu.log
