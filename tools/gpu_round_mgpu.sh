#!/bin/bash
# Multi-GPU round (run with gpurun --gpus N): multi-GPU parity tests, then bench.py: cfg2 at 2..N GPUs, cfg3 / cfg4 at N GPUs.
mkdir -p gpurun_out
TAG=${1:-r02}
NG=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_mgpu_${TAG}.log 2>&1
echo "pytest rc=$?"; grep -E "engine-side|DDP\+SyncBN|BatchNorm-bias|running stat|sharded sweep|MGPU_CHECK|passed|failed|skipped" gpurun_out/pytest_mgpu_${TAG}.log | grep -v "rank [1-7]/"
run_bench() {   # n config
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2957$1 bench.py \
      --gpus $1 --steps 10 --warmup 3 --no-fast-mode --config $2 > gpurun_out/bench_${TAG}_$2_n$1.json 2> gpurun_out/bench_${TAG}_$2_n$1.err
  echo "bench $2 n=$1 rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${TAG}_$2_n$1.json').read().strip().splitlines()[-1])
print('$2 n=$1', round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'sk', d['sk'].get('iters_per_sec'), 'sweep', (d['sweep'] or {}).get('value'), 'assign', d['assign'], 'incl', (d['incl_sk'] or {}).get('value'))
" || tail -5 gpurun_out/bench_${TAG}_$2_n$1.err
}
for n in 2 4 8; do
  if [ $n -le $NG ] && { [ $n -eq $NG ] || [ "$2" != "lastonly" ]; }; then run_bench $n cfg2; fi
done
run_bench $NG cfg3
run_bench $NG cfg4
