#!/bin/bash
# round 2, call H (1 GPU): 16-warp dK/dV kernel probe, full GPU suite on the final tree, bench
mkdir -p gpurun_out
timeout 120 python tests/dkdv16_probe.py > gpurun_out/r2h_dkdv16.log 2>&1; echo "probe rc=$?" >> gpurun_out/r2h_dkdv16.log
cut -c1-330 gpurun_out/r2h_dkdv16.log
timeout 400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -15 gpurun_out/r2h_pytest.log | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench1.log 2> gpurun_out/r2h_bench1.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/r2h_bench1.log; tail -3 gpurun_out/r2h_bench1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
