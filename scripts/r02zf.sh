#!/bin/bash
# session ZF: final validation at the last commit of the round -- smoke, every -m gpu test in ONE process (the way the
# round-end driver runs it), the default bench line, then a fresh ncu launch list of a short C4 bench run
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.log; }
tail -n 1 gpurun_out/smoke.log
timeout -k 10 200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r02zf_gpu_tests_single_process.log 2>&1
echo "gpu tests rc $?: $(tail -n 1 gpurun_out/r02zf_gpu_tests_single_process.log)"
grep -h "AssertionError\|^E  \|^FAILED" gpurun_out/r02zf_gpu_tests_single_process.log | cut -c1-300 | head -20
timeout 400 python bench.py > gpurun_out/r02zf_bench_c4.json 2> gpurun_out/bench_err.log; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zf_bench_c4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity']['ok'], repr(d['metrics']['mAP']), d['clocks'])
print({k:(v.get('ms_per_step'), v.get('fused_tflops'), v.get('queries_per_sec')) for k,v in d['secondary'].items()})
print({k:(v['us_per_step_graph']) for k,v in d['sdm'].items()})
PY
KERN='regex:retrieve_fused|rescore_topk|cand_select|pos_scores|pos_sort|l2norm_rows|mm_fuse|sim_gemm|calib_split|hist_to_above|metrics|pid_lookup|topk_check|merge_topk'
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file gpurun_out/r02zf_launches_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-sdm --no-cpu-baseline --no-secondary --parity-queries 0 > gpurun_out/ncu_c4.log 2>&1
echo "ncu rc $?"
python - <<'PY'
import csv, collections
try:
    rows = [r for r in csv.reader(open('gpurun_out/r02zf_launches_c4.csv')) if len(r) > 5]
    hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        a = agg.setdefault(r[ki][:60], [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(v)
    tot = sum(a[1] for a in agg.values())
    with open('gpurun_out/r02zf_launches_c4.txt', 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none of: python bench.py --workload c4 --steps 1 --warmup 3 --no-sdm --no-cpu-baseline --no-secondary --parity-queries 0\n')
        for k, (n, t, l) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            line = '%-62s n=%4d total %9.3f ms (%5.1f %%) avg %9.1f us  last launches (us): %s' % (k, n, t / 1e6, 100 * t / tot, t / n / 1e3, [round(x / 1e3) for x in l[-6:]])
            print(line); f.write(line + '\n')
except Exception as e:
    print('launch list not produced:', e)
PY
