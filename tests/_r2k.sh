#!/bin/bash
# round 2, call K (1 GPU, short): D = 64 forward knobs (P hand-off parts, FMA-pipe emulation share) A/B + parity of each
mkdir -p gpurun_out
: > gpurun_out/r2k_ab.log
for lib in ab/*.so; do
  echo "== $lib" >> gpurun_out/r2k_ab.log
  FA_B200_LIB=$PWD/$lib timeout 60 python tests/perf_probe.py d64 2>&1 | head -4 >> gpurun_out/r2k_ab.log
  FA_B200_LIB=$PWD/$lib timeout 90 python -m pytest tests/test_gpu_forward.py -m gpu -q -p no:cacheprovider -k "not fp32 and not flagship and not full" 2>&1 | tail -1 >> gpurun_out/r2k_ab.log
done
cat gpurun_out/r2k_ab.log
