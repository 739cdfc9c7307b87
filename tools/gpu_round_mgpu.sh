#!/bin/bash
# Multi-GPU round (run with gpurun --gpus N): multi-GPU parity tests, then bench.py at 2..N GPUs.
mkdir -p gpurun_out
TAG=${1:-r02}
NG=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_mgpu_${TAG}.log 2>&1
echo "pytest rc=$?"; grep -v "Randomy\|resnet9, dur\|Using MLP" gpurun_out/pytest_mgpu_${TAG}.log | tail -60
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py \
        --gpus $n --steps 10 --warmup 3 --no-fast-mode > gpurun_out/bench_${TAG}_n$n.json 2> gpurun_out/bench_${TAG}_n$n.err
    echo "bench n=$n rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${TAG}_n$n.json').read().strip().splitlines()[-1])
print('n=$n', round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'sk', d['sk'].get('iters_per_sec'), 'sweep', (d['sweep'] or {}).get('value'), 'assign', (d['assign'] or {}).get('seconds'))
" || tail -5 gpurun_out/bench_${TAG}_n$n.err
  fi
done
