#!/bin/bash
# ncu --set full captures of the tap-reuse conv kernel and the weight-gradient kernel on the layer-1 shapes; summaries in gpurun_out/
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -o gpurun_out/prof_halo_${TAG} -f \
    python tools/prof_halo.py > gpurun_out/ncu_halo_${TAG}.log 2>&1
ncu -i gpurun_out/prof_halo_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py \
    "fwd l1 1x3x3 64->144 (warm-up)" "fwd l1 1x3x3 64->144" "fwd l1 3x1x1 144->64 (warm-up)" "fwd l1 3x1x1 144->64" \
    "dgrad l1 1x3x3 (dz 144 -> dx 64) (warm-up)" "dgrad l1 1x3x3 (dz 144 -> dx 64)" "dgrad l1 3x1x1 (dz 64 -> dx 144) (warm-up)" \
    "dgrad l1 3x1x1 (dz 64 -> dx 144)" > gpurun_out/ncu_halo_summary_${TAG}.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_bf16_kernel -o gpurun_out/prof_wgrad_${TAG} -f \
    python tools/prof_wgrad.py > gpurun_out/ncu_wgrad_${TAG}.log 2>&1
ncu -i gpurun_out/prof_wgrad_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py \
    "wgrad l1 1x3x3 64->144 (warm-up)" "wgrad l1 1x3x3 64->144" "wgrad l1 3x1x1 144->64, exchanged operands (warm-up)" \
    "wgrad l1 3x1x1 144->64, exchanged operands" > gpurun_out/ncu_wgrad_summary_${TAG}.txt
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/ncu_halo_summary_${TAG}.txt gpurun_out/ncu_wgrad_summary_${TAG}.txt | grep -E "^##|time_duration|tensor_cycles|dram__bytes"
