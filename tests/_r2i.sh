#!/bin/bash
# round 2, call I (1 GPU, short): backward + ring suites on the final tree (after removing the 16-warp experiment)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_ring.py tests/test_gpu_harness.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r2i_pytest.log
