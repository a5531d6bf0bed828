#!/bin/bash
# session ZE: label form of the SDM entry points on the CUDA-core kernels (small / step / general): SDM kernel tests,
# the compute_loss fixtures (now all through the label form), C2 / C5 step times
mkdir -p gpurun_out
timeout -k 10 150 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "sdm" -p no:cacheprovider > gpurun_out/r02ze_test_sdm.log 2>&1
echo "sdm kernel tests rc $?: $(tail -n 1 gpurun_out/r02ze_test_sdm.log)"
timeout -k 10 100 python -m pytest tests/test_zz_protocol_gpu.py -q -m gpu -k "sdm or alignment" -p no:cacheprovider > gpurun_out/r02ze_test_sdm_proto.log 2>&1
echo "sdm fixture tests rc $?: $(tail -n 1 gpurun_out/r02ze_test_sdm_proto.log)"
grep -h "AssertionError\|^E  \|^FAILED\|Error" gpurun_out/r02ze_test_sdm*.log | cut -c1-240 | sort | uniq -c | sort -rn | head -20
timeout 90 python scripts/sdm_bench.py > gpurun_out/r02ze_sdm_bench.txt 2>&1; echo "sdm_bench rc $?"
cat gpurun_out/r02ze_sdm_bench.txt | tr -d '\n' | cut -c1-900
