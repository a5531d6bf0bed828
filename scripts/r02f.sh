#!/bin/bash
# round-2 GPU session F: all GPU tests, ncu evidence (launch list + --set full of every kernel), compute-sanitizer
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
TMO=900 bash scripts/gpu_tests_staged.sh 2>&1 | tail -12
grep -h "AssertionError\|^E  " gpurun_out/test_*.log | cut -c1-300 | sort | uniq -c | sort -rn | head -20
echo "=== default"; bash scripts/bench_short.sh c4 2>&1 | tee gpurun_out/ab_default.log
echo "=== c3b"; bash scripts/bench_short.sh c3b 2>&1 | tee gpurun_out/ab_c3b.log
KERN='regex:retrieve_fused|rescore_topk|cand_select|pos_scores|pos_sort|l2norm_rows|mm_fuse|sim_gemm|calib_split|hist_to_above|metrics|pid_lookup|topk_check|tc_prep|tc_fwd|tc_bwd|sdm_small'
echo "=== ncu launch list (bench c4, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file gpurun_out/r02f_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-sdm --no-cpu-baseline --no-secondary --parity-queries 0 > gpurun_out/ncu_c4.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02f_launches_c4.csv')) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    a = agg.setdefault(r[ki][:60], [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(v)
tot = sum(a[1] for a in agg.values())
with open('gpurun_out/r02f_launches_c4.txt', 'w') as f:
    for k, (n, t, l) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = '%-62s n=%4d total %9.3f ms (%5.1f %%) avg %9.1f us  last launches (us): %s' % (k, n, t / 1e6, 100 * t / tot, t / n / 1e3, [round(x / 1e3) for x in l[-6:]])
        print(line); f.write(line + '\n')
PY
echo "=== ncu --set full"
timeout 1500 ncu --set full --clock-control none --import-source on -k "$KERN" --launch-skip 0 -c 80 -o gpurun_out/r02f_full -f python scripts/ncu_targets.py > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/r02f_full.ncu-rep
echo "=== compute-sanitizer memcheck"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_targets.py > gpurun_out/r02f_memcheck.log 2>&1; tail -8 gpurun_out/r02f_memcheck.log
echo "=== compute-sanitizer racecheck (tiny)"
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_targets.py tiny > gpurun_out/r02f_racecheck.log 2>&1; tail -8 gpurun_out/r02f_racecheck.log
echo "=== compute-sanitizer synccheck (tiny)"
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_targets.py tiny > gpurun_out/r02f_synccheck.log 2>&1; tail -5 gpurun_out/r02f_synccheck.log
