#!/bin/bash
# session O: ncu --set full of the three SDM tcgen05 kernels at C5 shapes (4 pairs), summary + hot SASS; the report comes back too
mkdir -p gpurun_out /tmp/rep
timeout 300 python scripts/sdm_step_once.py 64 8 bf16 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_ --launch-skip 9 -c 3 -o /tmp/rep/sdm_c5 -f python scripts/sdm_step_once.py 64 8 bf16 > gpurun_out/ncu_full_sdm.log 2>&1; tail -2 gpurun_out/ncu_full_sdm.log
ls -la /tmp/rep
python scripts/ncu_summary.py /tmp/rep/sdm_c5.ncu-rep gpurun_out/r02o_sdm_tc_c5_ncu_full_summary.txt > /dev/null 2>&1
python scripts/ncu_hot_sass.py /tmp/rep/sdm_c5.ncu-rep tc_fwd 45 > gpurun_out/r02o_sdm_fwd_hot_sass.txt 2>&1
python scripts/ncu_hot_sass.py /tmp/rep/sdm_c5.ncu-rep tc_bwd 45 > gpurun_out/r02o_sdm_bwd_hot_sass.txt 2>&1
cp /tmp/rep/sdm_c5.ncu-rep gpurun_out/r02o_sdm_c5.ncu-rep
grep -A1 "Kernel Name" gpurun_out/r02o_sdm_tc_c5_ncu_full_summary.txt | grep -v "^--" | cut -c1-160
