#!/bin/bash
# session X (8 GPUs): N = 8 scaling line with single-launch gallery pass + warp-level cand_select; NCCL world-8 parity test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_nccl_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep -E "^\{|passed|failed|skipped|Error|assert" | cut -c1-1500 | tee gpurun_out/r02x_test_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 5 --warmup 3 --no-sdm --no-cpu-baseline --no-secondary --parity-queries 64 > gpurun_out/r02x_scale_n8.json 2> gpurun_out/r02x_scale_n8.err
tail -1 gpurun_out/r02x_scale_n8.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N=%d qps %.0f ms %.2f e2e_ms %.2f e2e_qps %.0f frac %.3f flagged %s mAP %.7f clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['run_info']['flagged_queries'], d['metrics']['mAP'], d['clocks']))
print('   kernels', d['kernel_ms_per_step']); print('   parity', {k: d['parity'][k] for k in ('d_mAP','cmc','cmc_oracle','cmc_rank_mismatches','top10_lists_differing_beyond_2e-6_ties','per_query_dAP_max','per_query_dAP_mean','ok')}); print('   e2e', d['e2e'])" || tail -5 gpurun_out/r02x_scale_n8.err
