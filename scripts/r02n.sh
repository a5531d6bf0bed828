#!/bin/bash
# session N: SDM tcgen05 path with the dense masks formed in the prep launch and 16 dS producer warps: SDM tests, step times, phase stamps
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -k sdm -q -m gpu -p no:cacheprovider > gpurun_out/test_sdm.log 2>&1; echo "sdm tests rc $?"; tail -3 gpurun_out/test_sdm.log
timeout -k 10 600 python -m pytest tests/test_zz_protocol_gpu.py -k "sdm or alignment" -q -m gpu -p no:cacheprovider > gpurun_out/test_sdm_proto.log 2>&1; echo "proto rc $?"; tail -2 gpurun_out/test_sdm_proto.log
grep -h "AssertionError\|^E  \|^FAILED" gpurun_out/test_sdm*.log | cut -c1-300 | sort | uniq -c | sort -rn | head
timeout 300 python scripts/sdm_bench.py 2>&1 | tail -30 | tee gpurun_out/r02n_sdm_bench.txt
REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_sdmtime.so timeout 300 python scripts/sdm_phase_times.py 2>&1 | tail -6 | tee gpurun_out/r02n_sdm_phases.txt
