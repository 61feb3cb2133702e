#!/bin/bash
# Round-end evidence pass on one B200: plain bench, ncu launch list of the same command, one
# `ncu --set full` capture of each flagship kernel, the harness sweep.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/p_bench_plain.log 2>&1 || exit 1
tail -c 1200 gpurun_out/p_bench_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/p_bench_ncu.log 2>&1
python tests/ncu_target.py bwd > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel|bwd_' -s 3 -c 4 -f -o gpurun_out/p_flagship \
    python tests/ncu_target.py bwd > gpurun_out/p_flagship_ncu.log 2>&1
tail -3 gpurun_out/p_flagship_ncu.log
(cd harness && ./flash_attn --dtype bf16 --csv ../gpurun_out/p_benchmark_results_bf16.csv > ../gpurun_out/p_harness_bf16.log 2>&1; tail -12 ../gpurun_out/p_harness_bf16.log)
