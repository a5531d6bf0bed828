#!/bin/bash
# Run the GPU parity tests in separate processes (a hung kernel then only loses its own group).
# Full logs land in gpurun_out/test_<group>.log; the summary line of each group in gpurun_out/test_summary.log.
mkdir -p gpurun_out
: > gpurun_out/test_summary.log
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; rc=$?; tail -3 gpurun_out/test_$name.log; echo "$name exit $rc: $(tail -1 gpurun_out/test_$name.log)" | tee -a gpurun_out/test_summary.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
run norm   tests/test_gpu_kernels.py -k "l2norm or fuse"
run sdm    tests/test_gpu_kernels.py -k "sdm"
run exact  tests/test_gpu_kernels.py -k "exact or pid_index"
run gemm   tests/test_gpu_kernels.py -k "sim_gemm"
run fused  tests/test_gpu_kernels.py -k "fused"
run rest   tests/test_gpu_kernels.py -k "not (l2norm or fuse or sdm or exact or pid_index or sim_gemm or fused)"
run proto  tests/test_zz_protocol_gpu.py
run native tests/test_zz_native_host_gpu.py
for f in tests/test_*gpu*.py; do case $f in tests/test_gpu_kernels.py|tests/test_zz_protocol_gpu.py|tests/test_zz_native_host_gpu.py) ;; *) run $(basename $f .py) $f ;; esac; done
cat gpurun_out/test_summary.log
