#!/bin/bash
# One GPU-box round (1 GPU): gpu test suite, smoke, bench (with per-layer detail), optional ncu launch list of one step.
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 1500 python -m pytest tests --maxfail=8 -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_${TAG}.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
SELAVI_BENCH_DETAIL=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
if [ "$2" == "ncu" ]; then
  SELAVI_BENCH_NO_SETTLE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5200 --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/ncu_bench_${TAG}.log 2>&1
  python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv --step 4 > gpurun_out/launch_list_summary_${TAG}.txt 2>&1
  head -40 gpurun_out/launch_list_summary_${TAG}.txt
fi
