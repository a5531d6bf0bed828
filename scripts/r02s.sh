#!/bin/bash
# session S: SDM backward with one fp16 dL/dS plane (A fp16 x B bf16): tests, step times A/B against the two-plane and the no-PDL builds,
# phase stamps; fused retrieval with deterministic calibration classes: fused / oracle tests, two bench lines (metrics must repeat)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -k sdm -q -m gpu -p no:cacheprovider > gpurun_out/test_sdm.log 2>&1; echo "sdm tests rc $?"; tail -3 gpurun_out/test_sdm.log
timeout -k 10 600 python -m pytest tests/test_zz_protocol_gpu.py -k "sdm or alignment" -q -m gpu -p no:cacheprovider > gpurun_out/test_sdm_proto.log 2>&1; echo "proto rc $?"; tail -2 gpurun_out/test_sdm_proto.log
grep -h "AssertionError\|^E  \|^FAILED" gpurun_out/test_sdm*.log | cut -c1-300 | sort | uniq -c | sort -rn | head
for v in default dsbf16 nopdl; do
  if [ $v = default ]; then unset REID_LIB; else export REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_$v.so; fi
  echo "== $v"; timeout 300 python scripts/sdm_bench.py 2>&1 | grep -A2 "c5_p64k8_bf16_10pairs\|c2_p4k2" | grep "graph\|pairs"
done 2>&1 | tee gpurun_out/r02s_sdm_ab.txt
unset REID_LIB
REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_sdmtime.so timeout 300 python scripts/sdm_phase_times.py 2>&1 | tail -6 | tee gpurun_out/r02s_sdm_phases.txt
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -k "host_query or fused" tests/test_fused_oracle_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do
timeout 900 python bench.py --no-sdm --no-secondary --no-cpu-baseline > gpurun_out/r02s_bench_c4_$i.json 2> gpurun_out/bench_err.log; echo "bench rc $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02s_bench_c4_$i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['kernel_ms_per_step']['reid_retrieve_fused'], d['roofline']['frac'], d['parity']['ok'], repr(d['metrics']['mAP']), d['clocks'])
PY
done
