#!/bin/bash
# A/B of an environment switch on the bench (no CPU baseline): usage gpu_ab.sh VAR valA valB
mkdir -p gpurun_out
for v in $2 $3; do
  env $1=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$1_$v.json 2> gpurun_out/ab_$1_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$1_$v.json").read().strip().splitlines()[-1])
print("$1=$v", "ms_per_step", round(d["ms_per_step"],2), "clips/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
PY
done
