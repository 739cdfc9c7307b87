#!/bin/bash
# ncu --set full capture of ONE train step of bench.py for the weight-gradient kernel set (the roofline kernel) and of the
# tap-reuse conv launches of the same step; raw pages -> profiles-ready summaries + ncu_traffic.json.  Outputs in gpurun_out/.
TAG=${1:-r02}
mkdir -p gpurun_out
export SELAVI_BENCH_NO_SETTLE=1
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode"
# step index 3 (0-based): 101 matching launches per step (49 wgrad + 49 reduce + 3 split)
timeout 420 ncu --set full --clock-control none -k regex:'wgrad_bf16_kernel|wgrad_reduce_kernel|split_bf16_kernel' -s 303 -c 101 \
    --csv --page raw --log-file gpurun_out/ncu_wgrad_step_${TAG}.csv $BENCH > gpurun_out/ncu_wgrad_step_${TAG}.log 2>&1
echo "ncu wgrad rc=$?"
python tools/ncu_traffic.py gpurun_out/ncu_wgrad_step_${TAG}.csv > gpurun_out/ncu_traffic_wgrad_${TAG}.json
python tools/ncu_summary.py < gpurun_out/ncu_wgrad_step_${TAG}.csv > gpurun_out/ncu_wgrad_step_summary_${TAG}.txt
if [ "$2" == "halo" ]; then
# tap-reuse conv launches of step 3: 57 per step (28 forward + 29 data gradient)
timeout 420 ncu --set full --clock-control none -k regex:conv_halo_kernel -s 171 -c 57 \
    --csv --page raw --log-file gpurun_out/ncu_halo_step_${TAG}.csv $BENCH > gpurun_out/ncu_halo_step_${TAG}.log 2>&1
echo "ncu halo rc=$?"
python tools/ncu_summary.py < gpurun_out/ncu_halo_step_${TAG}.csv > gpurun_out/ncu_halo_step_summary_${TAG}.txt
fi
cat gpurun_out/ncu_traffic_wgrad_${TAG}.json
grep -c "^## launch" gpurun_out/ncu_wgrad_step_summary_${TAG}.txt
