#!/bin/bash
# One GPU-box round: full gpu test suite, smoke, bench, ncu launch list + full captures.  Outputs in gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_${TAG}.log 2>&1
tail -5 gpurun_out/pytest_${TAG}.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json
if [ "$2" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1600 --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 40 -c 2 -o gpurun_out/prof_conv_${TAG} \
      python tools/quick_bench.py conv > gpurun_out/ncu_conv_${TAG}.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sk_kernel -s 3 -c 1 -o gpurun_out/prof_sk_${TAG} \
      python tools/quick_bench.py sk > gpurun_out/ncu_sk_${TAG}.log 2>&1
  ls -la gpurun_out/
fi
