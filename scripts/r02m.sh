#!/bin/bash
# session M: proportional calibration retirement + seeded candidate thresholds: fused / oracle / protocol tests, emulated 8-way shard pass
# (time + launch list), default bench line
mkdir -p gpurun_out
: > gpurun_out/test_summary.log
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; rc=$?; tail -3 gpurun_out/test_$name.log; echo "$name exit $rc: $(tail -1 gpurun_out/test_$name.log)" | tee -a gpurun_out/test_summary.log; }
run fused  tests/test_gpu_kernels.py -k "fused"
run oracle tests/test_fused_oracle_gpu.py
run proto  tests/test_zz_protocol_gpu.py
grep -h "AssertionError\|^E  \|^FAILED" gpurun_out/test_*.log | cut -c1-300 | sort | uniq -c | sort -rn | head -20
for wd in 8 2 1; do timeout 300 python scripts/shard_probe.py $wd c4 2>&1 | tail -2; done | tee gpurun_out/r02m_shard_probe.txt
timeout 300 python scripts/shard_probe.py 1 c3b 2>&1 | tail -2 | tee -a gpurun_out/r02m_shard_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"retrieve_fused|cand_select|calib_split|hist_to_above" -c 10 --csv --log-file gpurun_out/r02m_launches.csv python scripts/shard_probe.py 8 c4 > gpurun_out/r02m_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02m_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
for r in rows[1:]:
    print(r[ki][:60].ljust(60), r[vi])
PY
echo "=== bench default"
timeout 900 python bench.py --no-sdm > gpurun_out/r02m_bench_c4.json 2> gpurun_out/bench_err.log; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench_c4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['parity']['ok'], d['parity']['d_mAP'], d['parity']['per_query_dAP_max'], d['parity']['per_query_dAP_mean'], d['metrics'], d['clocks'], d['run_info'])
print({k:(v.get('ms_per_step'), v.get('fused_kernel_ms'), v.get('fused_tflops'), v.get('metrics',{}).get('mAP')) for k,v in d['secondary'].items()})
PY
