#!/bin/bash
# session ZH: sdm_alignment_loss with the masks stacked once and the :608-625 tail inside the autograd Function:
# the compute_loss fixtures + the SDM kernel tests on the GPU, then the eager step time again
mkdir -p gpurun_out
timeout -k 5 40 python -m pytest tests/test_zz_protocol_gpu.py tests/test_gpu_kernels.py -q -m gpu -k "sdm or alignment" -p no:cacheprovider > gpurun_out/r02zh_test_sdm.log 2>&1
echo "sdm tests rc $?: $(tail -n 1 gpurun_out/r02zh_test_sdm.log)"
grep -h "AssertionError\|^E  \|^FAILED\|Error" gpurun_out/r02zh_test_sdm.log | cut -c1-240 | head -10
timeout 25 python scripts/alignment_bench.py > gpurun_out/r02zh_alignment_bench.txt 2>&1; echo "bench rc $?"
cat gpurun_out/r02zh_alignment_bench.txt | tr -d "\n " | cut -c1-1000
