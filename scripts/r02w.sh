#!/bin/bash
# session W: per-kernel times of the gallery pass (cand_select warp + CTA fallback) on the emulated 8-way shard, the full shard and C3b
mkdir -p gpurun_out
for args in "8 c4" "1 c4" "1 c3b"; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"retrieve_fused|cand_select|calib_split|hist_to_above" --launch-skip 15 -c 5 --csv --log-file gpurun_out/r02w_launches.csv python scripts/shard_probe.py $args > gpurun_out/r02w_ncu.log 2>&1
echo "== shard_probe $args"; tail -1 gpurun_out/r02w_ncu.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02w_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
for r in rows[1:]:
    print(r[ki][:60].ljust(60), r[vi])
PY
done 2>&1 | tee gpurun_out/r02w_cand_select.txt
