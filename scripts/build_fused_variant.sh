#!/bin/bash
# build_fused_variant.sh <name> [extra nvcc flags]: libreid_b200 with retrieve_fused.cu compiled with extra -D switches
# (A/B experiments on the GPU box: REID_LIB=prcv2025reid_b200/variants/libreid_<name>.so python bench.py ...)
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $here/prcv2025reid_b200/variants /tmp/reid_variants
python -c "from prcv2025reid_b200 import build; build.build_library()" >/dev/null
obj=/tmp/reid_variants/retrieve_fused_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
     -c $here/prcv2025reid_b200/csrc/retrieve_fused.cu -o $obj
objs=""
for f in normalize pos_index rank sdm sdm_tc sim_gemm api; do objs="$objs $here/prcv2025reid_b200/build/$f.o"; done
nvcc -shared -o $here/prcv2025reid_b200/variants/libreid_$name.so $objs $obj
echo built prcv2025reid_b200/variants/libreid_$name.so
