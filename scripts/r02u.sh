#!/bin/bash
# session U (2 GPUs): NCCL parity test at world 2 after the deterministic metrics reduction
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_nccl_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep -E "^\{|passed|failed|skipped|Error|assert" | cut -c1-1500 | tee gpurun_out/r02u_test_nccl.log
