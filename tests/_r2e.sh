#!/bin/bash
# round 2, call E (8 GPUs): bench.py at N=8 (replicas + config-4 split + ring with parity), harness presets 4 and 5
# in one process, single-process group check, peer vs NCCL transport at medium N
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2e_clocks.csv 2>&1 &
SMI=$!
timeout 400 $TR --master-port 29651 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2e_bench8.log 2> gpurun_out/r2e_bench8.err
echo "bench8 rc=$?"; grep -o '"ring": .*' gpurun_out/r2e_bench8.log | cut -c1-3000; tail -3 gpurun_out/r2e_bench8.err
(cd harness && timeout 200 ./flash_attn --config 4 --gpus 8 > ../gpurun_out/r2e_harness_cfg4.log 2>&1; echo "rc=$?" >> ../gpurun_out/r2e_harness_cfg4.log)
cat gpurun_out/r2e_harness_cfg4.log
(cd harness && timeout 400 ./flash_attn --config 5 --gpus 8 > ../gpurun_out/r2e_harness_cfg5.log 2>&1; echo "rc=$?" >> ../gpurun_out/r2e_harness_cfg5.log)
cat gpurun_out/r2e_harness_cfg5.log
timeout 200 python tests/mgpu_check.py --n-total 16384 --heads 5 --causal 1 --gpus 8 > gpurun_out/r2e_mgpu.log 2>&1; echo "mgpu rc=$?" >> gpurun_out/r2e_mgpu.log
cut -c1-900 gpurun_out/r2e_mgpu.log
for tp in peer nccl gather; do
  timeout 200 $TR --master-port 29652 tests/ring_check.py --n-total 131072 --heads 4 --hdim 128 --causal 1 --check 0 --reps 5 --transport $tp 2>&1 | grep -E "ring_forward|Error|error" | tail -2 >> gpurun_out/r2e_ring_transports.jsonl
done
timeout 200 $TR --master-port 29653 tests/ring_check.py --n-total 131072 --heads 4 --hdim 128 --causal 1 --check 0 --reps 3 --bwd 1 --transport nccl 2>&1 | grep -E "ring_forward|Error|error" | tail -2 >> gpurun_out/r2e_ring_transports.jsonl
cat gpurun_out/r2e_ring_transports.jsonl
kill $SMI
