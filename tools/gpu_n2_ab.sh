#!/bin/bash
# 2-GPU A/B of the DDP bucket size and the SyncBN exchange path (no CPU baseline)
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 \
      bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/n2_$tag.json").read().strip().splitlines()[-1])
print("$tag", "ms_per_step", round(d["ms_per_step"],2), "clips/s", round(d["value"],1), "step_ms", d["step_ms"])
PY
}
run default SELAVI_X=0
run bucket1g SELAVI_DDP_BUCKET_MB=1024
run bn_nccl SELAVI_BN_EXCHANGE=nccl
