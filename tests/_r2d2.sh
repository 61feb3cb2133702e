#!/bin/bash
# round 2, call D2 (2 GPUs): single-process group (fa_mgpu) checks, multi-GPU pytest, harness presets 4 and 5
mkdir -p gpurun_out
timeout 200 python tests/mgpu_check.py --n-total 2048 --heads 5 --causal 1 --gpus 2 > gpurun_out/r2d2_mgpu.log 2>&1; echo "mgpu causal rc=$?" >> gpurun_out/r2d2_mgpu.log
timeout 200 python tests/mgpu_check.py --n-total 2048 --heads 5 --causal 0 --gpus 2 >> gpurun_out/r2d2_mgpu.log 2>&1; echo "mgpu non-causal rc=$?" >> gpurun_out/r2d2_mgpu.log
timeout 200 python tests/mgpu_check.py --n-total 65536 --heads 4 --causal 1 --gpus 2 --reps 3 >> gpurun_out/r2d2_mgpu.log 2>&1; echo "mgpu 64k rc=$?" >> gpurun_out/r2d2_mgpu.log
cut -c1-900 gpurun_out/r2d2_mgpu.log
(cd harness && timeout 300 ./flash_attn --config 4 --gpus 2 > ../gpurun_out/r2d2_harness_cfg4.log 2>&1; echo "rc=$?" >> ../gpurun_out/r2d2_harness_cfg4.log)
cat gpurun_out/r2d2_harness_cfg4.log
(cd harness && timeout 400 ./flash_attn --config 5 --gpus 2 --max-n 262144 > ../gpurun_out/r2d2_harness_cfg5.log 2>&1; echo "rc=$?" >> ../gpurun_out/r2d2_harness_cfg5.log)
cat gpurun_out/r2d2_harness_cfg5.log
timeout 700 python -m pytest tests/test_gpu_ring_multi.py -m gpu -q -p no:cacheprovider > gpurun_out/r2d2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d2_pytest.log; tail -30 gpurun_out/r2d2_pytest.log
