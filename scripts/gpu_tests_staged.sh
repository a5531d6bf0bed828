#!/bin/bash
# Run the GPU parity tests in separate processes (a hung kernel then only loses its own group).
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-420}" python -m pytest "$@" -q -m gpu -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/test_$name.log; echo "exit ${PIPESTATUS[0]}"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
run norm   tests/test_gpu_kernels.py -k "l2norm or fuse"
run sdm    tests/test_gpu_kernels.py -k "sdm"
run exact  tests/test_gpu_kernels.py -k "exact or pid_index"
run gemm   tests/test_gpu_kernels.py -k "sim_gemm"
run fused  tests/test_gpu_kernels.py -k "fused"
run rest   tests/test_gpu_kernels.py -k "not (l2norm or fuse or sdm or exact or pid_index or sim_gemm or fused)"
run proto  tests/test_zz_protocol_gpu.py
run native tests/test_zz_native_host_gpu.py
