#!/bin/bash
# session L: launch list (ncu, per-launch durations) of the emulated 8-way shard pass
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"retrieve_fused|cand_select|calib_split|hist_to_above" -c 40 --csv --log-file gpurun_out/r02l_launches.csv python scripts/shard_probe.py 8 c4 > gpurun_out/r02l_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02l_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
for r in rows[1:]:
    print(r[ki][:60].ljust(60), r[vi])
PY
