#!/bin/bash
# usage: bench_short.sh <workload> [extra bench args]; prints a compact summary
w=$1; shift
timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-sdm --no-cpu-baseline --no-secondary --parity-queries 0 "$@" 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('qps %.0f ms %.2f fused_ms %.2f TF %.1f frac %.3f flagged %s mAP %.7f e2e_ms %.1f clocks %s' % (d['value'], d['ms_per_step'], d['kernel_ms_per_step'].get('reid_retrieve_fused',0), d['roofline']['achieved'], d['roofline']['frac'], d['run_info']['flagged_queries'], d['metrics']['mAP'], d['e2e']['ms_per_step'], d['clocks'])); print('   ', {k:v for k,v in d['kernel_ms_per_step'].items() if v > 0.05}, {k: (v['ms'], round(v['frac_of_hbm_peak'], 3)) for k, v in d.get('hbm_kernels', {}).items()})"
