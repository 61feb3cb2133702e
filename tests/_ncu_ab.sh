#!/bin/bash
# Development helper: one `ncu --set full` capture of the flagship forward kernel per library in ab/*.so
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename $lib .so)
  FA_B200_LIB=$PWD/$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_tc -s 2 -c 1 \
     -f -o gpurun_out/ncu_$name python tests/ncu_target.py fwd > gpurun_out/ncu_$name.log 2>&1
  tail -2 gpurun_out/ncu_$name.log
done
