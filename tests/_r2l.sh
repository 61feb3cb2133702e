#!/bin/bash
# round 2, call L (2 GPUs, short): ring forward+backward on the final tree (after the merge-plan refactor)
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 tests/ring_check.py --n-total 4096 --heads 3 --hdim 128 --causal 1 --check 1 --reps 2 --bwd 1 --transport auto 2>&1 | grep -E "ring_forward|FAILED|Error" | cut -c1-700 | tee gpurun_out/r2l_ring.log
