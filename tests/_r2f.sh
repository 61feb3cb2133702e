#!/bin/bash
# round 2, call F (1 GPU): ncu evidence -- full-set capture of every kernel the round names, launch list of bench.py
mkdir -p gpurun_out
timeout 100 python tests/ncu_target_r2.py > gpurun_out/r2f_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none -k regex:'fwd_tc_kernel|bwd_|flash_attention_v2' -c 20 -f -o gpurun_out/r2_prof \
    python tests/ncu_target_r2.py > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras --no-sustained > gpurun_out/r2f_bench_plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras --no-sustained > gpurun_out/r2f_bench_ncu.log 2>&1
tail -2 gpurun_out/r2f_bench_ncu.log | cut -c1-300
timeout 120 python -m pytest tests/test_gpu_ring.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
ls -la gpurun_out/r2_prof.ncu-rep
