#!/bin/bash
# round 2, call J (1 GPU, short): source-level ncu capture of the D = 64 kernels (config 4) for the stall attribution
mkdir -p gpurun_out
timeout 60 python tests/ncu_target_d64.py > gpurun_out/r2j_plain.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel|bwd_dkdv|bwd_dq' -s 3 -c 3 -f -o gpurun_out/r2_prof_d64 \
    python tests/ncu_target_d64.py > gpurun_out/r2j_ncu.log 2>&1
tail -3 gpurun_out/r2j_ncu.log; ls -la gpurun_out/r2_prof_d64.ncu-rep
