#!/bin/bash
# Final verification round (1 GPU): full gpu test suite, the cp.async fallback of the weight gradient, smoke, bench, and a
# metrics-only ncu pass over every weight-gradient launch of one bench step (DRAM bytes for roofline.traffic).
mkdir -p gpurun_out
TAG=${1:-r02z}
timeout 200 python -m pytest tests --maxfail=5 -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_${TAG}.log
SELAVI_WGRAD_TMA=0 timeout 80 python -m pytest tests/test_conv_gpu.py -k wgrad -q -m gpu -p no:cacheprovider 2>&1 | tail -2
timeout 80 python __graft_entry__.py smoke 2>&1 | tail -1
SELAVI_BENCH_DETAIL=1 timeout 260 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/bench_${TAG}.json'))
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel'], round(d['roofline']['frac'],3), 'sk', round(d['sk']['iters_per_sec']), 'sweep', round(d['sweep']['value']), 'lib', d['library_baseline'] and {k: round(v['value'],1) for k, v in d['library_baseline'].items() if isinstance(v, dict)})
"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
SELAVI_BENCH_NO_SETTLE=1 timeout 220 ncu --metrics $M --clock-control none -k regex:'wgrad_bf16_kernel|wgrad_reduce_kernel|split_bf16_kernel' -s 303 -c 101 \
    --csv --log-file gpurun_out/ncu_wgrad_metrics_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/ncu_wgrad_metrics_${TAG}.log 2>&1
echo "ncu rc=$?"; python tools/ncu_traffic.py gpurun_out/ncu_wgrad_metrics_${TAG}.csv | head -12
