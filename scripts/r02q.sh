#!/bin/bash
# session Q: host-query pipeline with the ramped block schedule: its GPU test and the e2e line of the bench
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -k "host_query or fused" -q -m gpu -p no:cacheprovider 2>&1 | tail -2
timeout 900 python bench.py --no-sdm --no-secondary --no-cpu-baseline > gpurun_out/r02q_bench_c4.json 2> gpurun_out/bench_err.log; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench_c4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['kernel_ms_per_step'], d['roofline']['frac'], d['parity']['ok'], d['clocks'])
PY
