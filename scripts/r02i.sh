#!/bin/bash
# round-2 GPU session I: ncu --set full of every retrieval kernel at HEAD, summarised ON the box (only text comes back)
mkdir -p gpurun_out /tmp/rep
KR='regex:retrieve_fused|rescore_topk|cand_select|pos_scores|pos_sort|l2norm_rows|mm_fuse|sim_gemm|calib_split|hist_to_above|metrics_kernel|pid_lookup|topk_check|merge_topk'
timeout 300 python scripts/ncu_targets.py retrieval > gpurun_out/ncu_targets_plain.log 2>&1 || { echo "TARGET FAILED"; tail -20 gpurun_out/ncu_targets_plain.log; exit 1; }
tail -1 gpurun_out/ncu_targets_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 14 -c 18 -o /tmp/rep/retrieval -f python scripts/ncu_targets.py retrieval > gpurun_out/ncu_full_retrieval.log 2>&1; tail -2 gpurun_out/ncu_full_retrieval.log
ls -la /tmp/rep
python scripts/ncu_summary.py /tmp/rep/retrieval.ncu-rep gpurun_out/r02i_retrieval_ncu_full_summary.txt > /dev/null 2>&1
python scripts/ncu_hot_sass.py /tmp/rep/retrieval.ncu-rep retrieve_fused 40 > gpurun_out/r02i_fused_hot_sass.txt 2>&1
grep -c "Kernel Name" gpurun_out/r02i_retrieval_ncu_full_summary.txt
grep -A1 "Kernel Name" gpurun_out/r02i_retrieval_ncu_full_summary.txt | grep -v "^--" | cut -c1-160
