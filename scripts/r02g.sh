#!/bin/bash
# round-2 GPU session G: ncu --set full of every kernel, summarised ON the box (only text comes back), tests of the last changes
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "store or merge or sim_gemm or label_form or train_eval" 2>&1 | tail -3
KR='regex:retrieve_fused|rescore_topk|cand_select|pos_scores|pos_sort|l2norm_rows|mm_fuse|sim_gemm|calib_split|hist_to_above|metrics_kernel|pid_lookup|topk_check'
KS='regex:tc_prep|tc_fwd|tc_bwd|sdm_small'
mkdir -p /tmp/rep
timeout 1500 ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 14 -c 16 -o /tmp/rep/retrieval -f python scripts/ncu_targets.py retrieval > gpurun_out/ncu_full_retrieval.log 2>&1; tail -2 gpurun_out/ncu_full_retrieval.log
timeout 900 ncu --set full --clock-control none --import-source on -k "$KS" -c 20 -o /tmp/rep/sdm -f python scripts/ncu_targets.py sdm > gpurun_out/ncu_full_sdm.log 2>&1; tail -2 gpurun_out/ncu_full_sdm.log
ls -la /tmp/rep
python scripts/ncu_summary.py /tmp/rep/retrieval.ncu-rep gpurun_out/r02g_retrieval_ncu_full_summary.txt > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/rep/sdm.ncu-rep gpurun_out/r02g_sdm_ncu_full_summary.txt > /dev/null 2>&1
python scripts/ncu_hot_sass.py /tmp/rep/retrieval.ncu-rep retrieve_fused 40 > gpurun_out/r02g_fused_hot_sass.txt 2>&1
grep -c "Kernel Name" gpurun_out/r02g_retrieval_ncu_full_summary.txt gpurun_out/r02g_sdm_ncu_full_summary.txt
head -30 gpurun_out/r02g_fused_hot_sass.txt
echo "=== c3b breakdown"; bash scripts/bench_short.sh c3b 2>&1 | tail -2
