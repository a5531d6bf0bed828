#!/bin/bash
# round-2 GPU session P (8 GPUs): NCCL parity tests (world 2 and 8), scaling runs N = 8, then N = 4 and N = 2 side by side on disjoint GPUs
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_zz_nccl_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep -E "^\{|passed|failed|skipped|Error|assert" | cut -c1-1200 | tee gpurun_out/r02p_test_nccl.log
show() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N=%d qps %.0f ms %.2f e2e_ms %.2f e2e_qps %.0f frac %.3f flagged %s mAP %.7f clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['run_info']['flagged_queries'], d['metrics']['mAP'], d['clocks']))
print('   kernels', d['kernel_ms_per_step']); print('   parity', {k: d['parity'][k] for k in ('d_mAP','cmc','cmc_oracle','cmc_rank_mismatches','top10_lists_differing_beyond_2e-6_ties','per_query_dAP_max','per_query_dAP_mean','ok')})" || tail -5 ${1%.json}.err; }
runb() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 5 --warmup 3 --no-sdm --no-cpu-baseline --no-secondary --parity-queries 64 > gpurun_out/r02p_scale_n$n.json 2> gpurun_out/r02p_scale_n$n.err; }
echo "=== N=8"; runb 8; show gpurun_out/r02p_scale_n8.json
echo "=== N=4 (GPUs 0-3) and N=2 (GPUs 4-5) side by side"
CUDA_VISIBLE_DEVICES=0,1,2,3 runb 4 &
CUDA_VISIBLE_DEVICES=4,5 runb 2 &
wait
show gpurun_out/r02p_scale_n4.json; show gpurun_out/r02p_scale_n2.json
