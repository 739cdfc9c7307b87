#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k halo -p no:cacheprovider -s > gpurun_out/pytest_halo.log 2>&1
tail -30 gpurun_out/pytest_halo.log
timeout 300 python tools/quick_bench.py halo > gpurun_out/qb_halo.log 2>&1
cat gpurun_out/qb_halo.log
