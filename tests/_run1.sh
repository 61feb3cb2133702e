mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1.log 2>&1
nproc >> gpurun_out/r1.log
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q -k "fp32" -x 2>&1 | tail -15 >> gpurun_out/r1.log
for c in "128 64 1 0" "128 64 1 0 vid" "128 64 0 0" "256 64 1 0" "384 64 1 1" "1000 64 1 0" "128 128 1 0" "128 128 1 0 vid" "1024 128 1 1" "2048 128 0 0"; do
  timeout 120 python tests/tc_probe.py $c >> gpurun_out/r1.log 2>&1 || echo "probe $c exit $?" >> gpurun_out/r1.log
done
tail -60 gpurun_out/r1.log
