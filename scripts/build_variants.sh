#!/bin/bash
# Builds variants of the library that differ in -D flags of merge_batch.cu (A/B runs in ONE gpurun call):
#   scripts/build_variants.sh name1 "-DIC_ROWS_DIST=1" name2 "-DIC_ROWS_DIST=3" ...
# -> build_variants/lib_<name>.so (git-ignored, shipped by gpurun); select with IMAGECLUST_B200_LIB=...
set -e
cd "$(dirname "$0")/.."
python -m imageclust_b200.build > /dev/null
mkdir -p build_variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
while [ $# -ge 2 ]; do
  name="$1"; defs="$2"; shift 2
  nvcc $FLAGS $defs -c imageclust_b200/csrc/merge_batch.cu -o build_variants/merge_batch_$name.o
  objs=$(ls imageclust_b200/build/*.o | grep -v merge_batch.o)
  nvcc -shared -o build_variants/lib_$name.so $objs build_variants/merge_batch_$name.o -cudart static -ldl
  echo "build_variants/lib_$name.so"
done
