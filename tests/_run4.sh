mkdir -p gpurun_out; : > gpurun_out/r4.log
for c in "128 64 1 0" "128 64 1 1" "64 64 0 0" "256 64 1 1" "1000 64 1 0" "1000 64 0 1" "128 128 1 0" "512 128 1 1" "2048 128 1 1" "777 128 0 0"; do
  timeout 180 python tests/bwd_probe.py $c >> gpurun_out/r4.log 2>&1 || echo "probe $c exit $?" >> gpurun_out/r4.log
done
cat gpurun_out/r4.log | tail -60
