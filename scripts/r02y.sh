#!/bin/bash
# session Y: ncu --set full of the kernels that changed since session I (fused calibration + main pass, calib_split, cand_select warp / CTA) at HEAD
mkdir -p gpurun_out /tmp/rep
KR='regex:retrieve_fused|cand_select|calib_split'
timeout 600 ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 5 -c 5 -o /tmp/rep/fused_head -f python scripts/ncu_targets.py retrieval > gpurun_out/ncu_full_fused_head.log 2>&1; tail -2 gpurun_out/ncu_full_fused_head.log
python scripts/ncu_summary.py /tmp/rep/fused_head.ncu-rep gpurun_out/r02y_fused_ncu_full_summary.txt > /dev/null 2>&1
python scripts/ncu_hot_sass.py /tmp/rep/fused_head.ncu-rep retrieve_fused 40 > gpurun_out/r02y_fused_hot_sass.txt 2>&1
grep -A1 "Kernel Name" gpurun_out/r02y_fused_ncu_full_summary.txt | grep -v "^--" | cut -c1-150
grep "pipe_tensor_cycles_active\|dram__bytes" gpurun_out/r02y_fused_ncu_full_summary.txt | cut -c1-130
