#!/bin/bash
# session V: warp-level cand_select + adaptive resident blocks: smoke, ALL gpu tests (staged), emulated 8-way shard pass, default bench line
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
TMO=420 bash scripts/gpu_tests_staged.sh 2>&1 | tail -14
grep -h "AssertionError\|^E  \|^FAILED" gpurun_out/test_*.log | cut -c1-300 | sort | uniq -c | sort -rn | head -20
timeout 200 python scripts/shard_probe.py 8 c4 2>&1 | tail -2 | tee gpurun_out/r02v_shard_probe.txt
echo "=== bench default"
timeout 900 python bench.py > gpurun_out/r02v_bench_c4.json 2> gpurun_out/bench_err.log; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v_bench_c4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['kernel_ms_per_step'], d['roofline']['frac'], d['parity']['ok'], repr(d['metrics']['mAP']), d['clocks'])
print({k:(v.get('ms_per_step'), v.get('fused_kernel_ms'), v.get('fused_tflops'), v.get('queries_per_sec')) for k,v in d['secondary'].items()})
print({k:(v['us_per_step_graph']) for k,v in d['sdm'].items()})
PY
