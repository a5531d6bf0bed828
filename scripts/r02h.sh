#!/bin/bash
# round-2 GPU session H: smoke, all GPU tests at HEAD (staged), the default bench line, memcheck of small cases
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
TMO=900 bash scripts/gpu_tests_staged.sh 2>&1 | tail -14
grep -h "AssertionError\|^E  \|^FAILED" gpurun_out/test_*.log | cut -c1-300 | sort | uniq -c | sort -rn | head -20
echo "=== bench default"
timeout 900 python bench.py > gpurun_out/r02h_bench_c4.json 2> gpurun_out/bench_err.log; echo "bench rc $?"; tail -c 3000 gpurun_out/r02h_bench_c4.json
echo "=== memcheck tiny"
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_targets.py tiny > gpurun_out/r02h_memcheck_tiny.log 2>&1; tail -6 gpurun_out/r02h_memcheck_tiny.log
