#!/bin/bash
# session K: per-rank gallery pass of an 8-way / 2-way sharded C4 run emulated on one GPU, library of HEAD vs work tree; SDM phase stamps
mkdir -p gpurun_out
for wd in 8 2; do
for lib in head work; do
  if [ $lib = head ]; then export REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_head.so; else unset REID_LIB; fi
  timeout 300 python scripts/shard_probe.py $wd c4 2>&1 | tail -2
done; done | tee gpurun_out/r02k_shard_probe.txt
unset REID_LIB
REID_LIB=$PWD/prcv2025reid_b200/variants/libreid_sdmtime.so timeout 300 python scripts/sdm_phase_times.py 2>&1 | tail -8 | tee gpurun_out/r02k_sdm_phases.txt
