#!/bin/bash
# like _ab.sh but one repetition and only the flagship lines
mode=${1:-bwd}
for lib in ab/*.so; do
  echo "== $lib"
  FA_B200_LIB=$PWD/$lib timeout 300 python tests/perf_probe.py $mode 2>&1 | grep -E "N=16384 d=128|d=64"
done
