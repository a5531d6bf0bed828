#!/bin/bash
# round-2 GPU session D: all GPU tests at the new re-scorer, calibration / push variants, launch list of C4
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
TMO=900 bash scripts/gpu_tests_staged.sh 2>&1 | tail -14
grep -h "^\[\|grad\|dq:\|dv:" gpurun_out/test_*.log | cut -c1-250 | head -80
echo "=== default"; bash scripts/bench_short.sh c4 2>&1 | tee gpurun_out/ab_default.log
for v in slow2 cal2k nohits; do
  echo "=== variant $v"; REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_$v.so bash scripts/bench_short.sh c4 2>&1 | tee gpurun_out/ab_$v.log
done
echo "=== c3b default"; bash scripts/bench_short.sh c3b 2>&1 | tee gpurun_out/ab_c3b_default.log
echo "=== ncu launch list c4"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-sdm --no-cpu-baseline --no-secondary --parity-queries 0 > gpurun_out/ncu_c4.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_c4.csv')) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    a = agg.setdefault(r[ki][:60], [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(v)
for k, (n, t, l) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]: print('%-62s n=%4d total %.3f ms avg %.1f us  last: %s' % (k, n, t / 1e6, t / n / 1e3, [round(x/1e3) for x in l[-6:]]))
PY
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 1500 gpurun_out/r02d_bench.json; tail -3 gpurun_out/r02d_bench.err
