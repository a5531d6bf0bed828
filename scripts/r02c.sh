#!/bin/bash
# round-2 GPU session C: oracle parity tests, experiment builds of the fused kernel, launch lists
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_fused_oracle_gpu.py tests/test_gpu_kernels.py -q -m gpu -k "oracle or fused or ragged or exact_ap or retrieve or c1" -p no:cacheprovider -s 2>&1 | grep -E "^\[|passed|failed|Error|assert" | cut -c1-260 | tee gpurun_out/test_oracle.log
echo "=== default"; bash scripts/bench_short.sh c4 2>&1 | tee gpurun_out/ab_default.log
for v in nohits main noload mmaonly st2 st4 hg3 q64; do
  echo "=== variant $v"; REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_$v.so bash scripts/bench_short.sh c4 2>&1 | tee gpurun_out/ab_$v.log
done
echo "=== c3b default"; bash scripts/bench_short.sh c3b 2>&1 | tee gpurun_out/ab_c3b_default.log
echo "=== c3b st4"; REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_st4.so bash scripts/bench_short.sh c3b 2>&1 | tee gpurun_out/ab_c3b_st4.log
echo "=== dbg counters"; REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_dbg.so timeout 300 python scripts/dbg_counters.py 2>&1 | tail -2 | tee gpurun_out/dbg.log
echo "=== ncu launch list c3b"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3b.csv python bench.py --workload c3b --steps 2 --warmup 3 --no-sdm --no-cpu-baseline --no-secondary --parity-queries 0 > gpurun_out/ncu_c3b.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_c3b.csv')) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    a = agg.setdefault(r[ki][:60], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]: print('%-62s n=%4d total %.3f ms avg %.1f us' % (k, n, t / 1e6, t / n / 1e3))
PY
