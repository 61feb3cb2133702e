#!/bin/bash
# round 2, call A (2 GPUs): full GPU test suite incl. the multi-GPU ring tests, bench at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_env.log 2>&1
nvidia-smi topo -m >> gpurun_out/r2a_env.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -40 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench1.log 2> gpurun_out/r2a_bench1.err
echo "bench1 rc=$?"; tail -c 3000 gpurun_out/r2a_bench1.log; tail -5 gpurun_out/r2a_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench2.log 2> gpurun_out/r2a_bench2.err
echo "bench2 rc=$?"; tail -c 6000 gpurun_out/r2a_bench2.log; tail -20 gpurun_out/r2a_bench2.err
