#!/bin/bash
# round-2 GPU session J: seeded candidate thresholds + per-call exchanges: smoke, fused / oracle / protocol tests, default bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
: > gpurun_out/test_summary.log
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; rc=$?; tail -3 gpurun_out/test_$name.log; echo "$name exit $rc: $(tail -1 gpurun_out/test_$name.log)" | tee -a gpurun_out/test_summary.log; }
run fused  tests/test_gpu_kernels.py -k "fused"
run oracle tests/test_fused_oracle_gpu.py
run proto  tests/test_zz_protocol_gpu.py
grep -h "AssertionError\|^E  \|^FAILED" gpurun_out/test_*.log | cut -c1-300 | sort | uniq -c | sort -rn | head -20
echo "=== bench default"
timeout 900 python bench.py --no-sdm > gpurun_out/r02j_bench_c4.json 2> gpurun_out/bench_err.log; echo "bench rc $?"; tail -c 2500 gpurun_out/r02j_bench_c4.json
