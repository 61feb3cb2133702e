#!/bin/bash
# round 2, call G (2 GPUs, short): the single-process group after the per-device enqueue threads, with deep queues
mkdir -p gpurun_out
timeout 90 python tests/mgpu_check.py --n-total 4096 --heads 4 --causal 1 --gpus 2 --reps 60 > gpurun_out/r2g_mgpu.log 2>&1; echo "mgpu deep-queue rc=$?" >> gpurun_out/r2g_mgpu.log
cut -c1-700 gpurun_out/r2g_mgpu.log
(cd harness && timeout 60 ./flash_attn --config 5 --gpus 2 --max-n 131072 > ../gpurun_out/r2g_harness_cfg5.log 2>&1; echo "rc=$?" >> ../gpurun_out/r2g_harness_cfg5.log)
cat gpurun_out/r2g_harness_cfg5.log
