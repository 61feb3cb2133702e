#!/bin/bash
# Development helper: time every library under ab/*.so with the same probe, same box.
# usage: tests/_ab.sh [fwd|bwd] ; results -> gpurun_out/ab_<mode>.log
mode=${1:-fwd}
mkdir -p gpurun_out
: > gpurun_out/ab_$mode.log
for rep in 1 2; do
for lib in ab/*.so; do
  echo "== $lib (rep $rep)" >> gpurun_out/ab_$mode.log
  if [ "$mode" = d64 ]; then FA_B200_LIB=$PWD/$lib timeout 300 python tests/perf_probe.py d64 >> gpurun_out/ab_$mode.log 2>&1
  elif [ "$mode" = bwd ]; then FA_B200_LIB=$PWD/$lib timeout 300 python tests/perf_probe.py bwd >> gpurun_out/ab_$mode.log 2>&1
  else FA_B200_LIB=$PWD/$lib timeout 300 python tests/perf_probe.py >> gpurun_out/ab_$mode.log 2>&1; fi
done
done
cat gpurun_out/ab_$mode.log
