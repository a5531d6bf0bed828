#!/bin/bash
# round-2 GPU session E (8 GPUs): NCCL parity tests (world 2 and 8), scaling runs N = 8, 4, 2, 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_zz_nccl_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep -E "^\{|passed|failed|skipped|Error|assert" | cut -c1-1200 | tee gpurun_out/test_nccl.log
for n in 8 4 2; do
  echo "=== N=$n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 5 --warmup 3 --no-sdm --no-cpu-baseline --no-secondary > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  tail -1 gpurun_out/scale_n$n.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N=%d qps %.0f ms %.2f e2e_ms %.2f frac %.3f flagged %s mAP %.7f clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['run_info']['flagged_queries'], d['metrics']['mAP'], d['clocks']))
print('   kernels', d['kernel_ms_per_step']); print('   parity', {k: d['parity'][k] for k in ('d_mAP','cmc','cmc_oracle','cmc_rank_mismatches','top10_lists_differing_beyond_2e-6_ties','per_query_dAP_max','per_query_dAP_mean','ok')})" || tail -5 gpurun_out/scale_n$n.err
done
echo "=== N=1"; bash scripts/bench_short.sh c4 2>&1 | tee gpurun_out/scale_n1.log
