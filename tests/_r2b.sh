#!/bin/bash
# round 2, call B (2 GPUs): fused backward first contact + A/B, ragged ring check with output, mgpu check
mkdir -p gpurun_out
timeout 300 python tests/bwd_ab.py > gpurun_out/r2b_bwd_ab.log 2>&1; echo "bwd_ab rc=$?" >> gpurun_out/r2b_bwd_ab.log
cat gpurun_out/r2b_bwd_ab.log
timeout 120 python tests/v2_probe.py > gpurun_out/r2b_v2.log 2>&1; cat gpurun_out/r2b_v2.log
timeout 700 python -m pytest tests/test_gpu_backward.py tests/test_gpu_ring.py tests/test_gpu_forward.py -m gpu -q --maxfail=8 -p no:cacheprovider -x > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log; tail -30 gpurun_out/r2b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29641 tests/ring_check.py --n-total 656 --heads 2 --hdim 64 --causal 1 --check 1 --reps 1 --bwd 1 --transport peer > gpurun_out/r2b_ragged.log 2>&1
echo "ragged rc=$?" >> gpurun_out/r2b_ragged.log; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2b_ragged.log | tail -15
timeout 200 $TR --master-port 29642 tests/ring_check.py --n-total 656 --heads 2 --hdim 64 --causal 1 --check 1 --reps 1 --bwd 0 --transport nccl > gpurun_out/r2b_ragged_nccl.log 2>&1
echo "ragged nccl rc=$?" >> gpurun_out/r2b_ragged_nccl.log; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2b_ragged_nccl.log | tail -8
timeout 200 python tests/mgpu_check.py --n-total 2048 --heads 5 --causal 1 --gpus 2 > gpurun_out/r2b_mgpu.log 2>&1
echo "mgpu rc=$?" >> gpurun_out/r2b_mgpu.log; tail -8 gpurun_out/r2b_mgpu.log
timeout 200 python tests/mgpu_check.py --n-total 2048 --heads 5 --causal 0 --gpus 2 >> gpurun_out/r2b_mgpu.log 2>&1
echo "mgpu nc rc=$?" >> gpurun_out/r2b_mgpu.log; tail -4 gpurun_out/r2b_mgpu.log
