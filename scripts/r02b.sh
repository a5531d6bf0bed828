#!/bin/bash
# round-2 GPU session B: smoke, all GPU tests, A/B of the fused-kernel variants, debug counters, bench
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
TMO=900 bash scripts/gpu_tests_staged.sh
grep -h "^\[" gpurun_out/test_test_fused_oracle_gpu.log | head -40
for v in base tree pf32 pf128 sf cal8k pf64sf; do
  echo "=== variant $v"; REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_$v.so bash scripts/bench_short.sh c4 2>&1 | tee gpurun_out/ab_$v.log
done
echo "=== dbg counters"; REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_dbg.so timeout 300 python scripts/dbg_counters.py 2>&1 | tail -2 | tee gpurun_out/dbg.log
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 2500 gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
