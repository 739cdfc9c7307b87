#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/umma_probe_shift.py > gpurun_out/probe_shift.log 2>&1
cat gpurun_out/probe_shift.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err
cat gpurun_out/bench_r01f.json
