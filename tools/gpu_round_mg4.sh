mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_mgpu_r02m.log 2>&1
echo "pytest rc=$?"; grep -E "engine-side|DDP\+SyncBN|BatchNorm-bias|running stat|sharded sweep|MGPU_CHECK|passed|failed|skipped" gpurun_out/pytest_mgpu_r02m.log | grep -v "rank [1-7]/"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 4 --steps 10 --warmup 3 --no-fast-mode > gpurun_out/bench_r02m_cfg2_n4.json 2> gpurun_out/bench_r02m_cfg2_n4.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02m_cfg2_n4.json').read().strip().splitlines()[-1])
print('n=4', round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'sk', d['sk'].get('iters_per_sec'), 'sweep', (d['sweep'] or {}).get('value'), 'assign', (d['assign'] or {}).get('seconds'), 'incl', (d['incl_sk'] or {}).get('value'))
" || tail -5 gpurun_out/bench_r02m_cfg2_n4.err
