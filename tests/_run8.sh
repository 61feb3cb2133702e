mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8; do
  timeout 300 $TR --nproc-per-node $n --master-port 29601 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu 2>&1 | grep '"metric"' > gpurun_out/r18_bench_${n}gpu.log
  cut -c1-400 gpurun_out/r18_bench_${n}gpu.log
done
: > gpurun_out/r18_ring8.log
for cfg in "--n-total 131072 --heads 4 --hdim 128 --causal 1 --check 0 --reps 3" "--n-total 131072 --heads 4 --hdim 128 --causal 0 --check 0 --reps 3" "--n-total 1048576 --heads 1 --hdim 128 --causal 1 --check 0 --reps 2" "--n-total 32768 --heads 4 --hdim 128 --causal 1 --check 1 --reps 3"; do
  timeout 300 $TR --nproc-per-node 8 --master-port 29602 tests/ring_check.py $cfg 2>&1 | grep -E "ring_forward|Error|error" | tail -2 >> gpurun_out/r18_ring8.log
done
cat gpurun_out/r18_ring8.log
# ring forward + backward, checked against the single-GPU kernels
timeout 300 $TR --nproc-per-node 8 --master-port 29603 tests/ring_check.py --n-total 65536 --heads 8 --hdim 128 --causal 1 --check 1 --reps 3 --bwd 1 2>&1 | grep -E "ring_|Error|error" | tail -3 >> gpurun_out/r18_ring8.log
tail -2 gpurun_out/r18_ring8.log
