#!/bin/bash
# session T: SDM backward with fp16 dL/dS plane x fp16 transposed operand image: tests, step times A/B (two bf16 planes / no PDL), phase stamps
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -k sdm -q -m gpu -p no:cacheprovider > gpurun_out/test_sdm.log 2>&1; echo "sdm tests rc $?"; tail -3 gpurun_out/test_sdm.log
timeout -k 10 600 python -m pytest tests/test_zz_protocol_gpu.py -k "sdm or alignment" -q -m gpu -p no:cacheprovider > gpurun_out/test_sdm_proto.log 2>&1; echo "proto rc $?"; tail -2 gpurun_out/test_sdm_proto.log
grep -h "AssertionError\|^E  \|^FAILED" gpurun_out/test_sdm*.log | cut -c1-300 | sort | uniq -c | sort -rn | head
for v in default dsbf16 nopdl default; do
  if [ $v = default ]; then unset REID_LIB; else export REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_$v.so; fi
  echo "== $v"; timeout 300 python scripts/sdm_bench.py 2>&1 | grep -A2 "c5_p64k8_bf16_10pairs\|c2_p4k2" | grep "graph\|pairs"
done 2>&1 | tee gpurun_out/r02t_sdm_ab.txt
unset REID_LIB
REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_sdmtime.so timeout 300 python scripts/sdm_phase_times.py 2>&1 | tail -6 | tee gpurun_out/r02t_sdm_phases.txt
