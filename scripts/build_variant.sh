#!/bin/bash
# build_variant.sh <git-ref|WORK> <out.so> [extra nvcc flags]: compile csrc of a git ref (or the work tree) into <out.so>
ref=$1; out=$2; shift 2
tmp=$(mktemp -d)
if [ "$ref" = "WORK" ]; then cp -r prcv2025reid_b200/csrc include $tmp/; mkdir -p $tmp/p; mv $tmp/csrc $tmp/p/; 
else git archive $ref prcv2025reid_b200/csrc include | tar -x -C $tmp; mkdir -p $tmp/p; mv $tmp/prcv2025reid_b200/csrc $tmp/p/; fi
mkdir -p $tmp/p/x; mv $tmp/p/csrc $tmp/p/x/csrc; mkdir -p $tmp/p/include; cp -r $tmp/include/* $tmp/p/include/   # csrc includes ../../include
objs=""
for f in normalize pos_index rank sdm sdm_tc sim_gemm retrieve_fused api; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $tmp/p/x/csrc/$f.cu -o $tmp/$f.o & 
  objs="$objs $tmp/$f.o"
done
wait
nvcc -shared -o $out $objs && echo built $out
rm -rf $tmp
