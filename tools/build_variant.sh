#!/bin/bash
# Variant build of the production library with extra -D flags (experiments): tools/build_variant.sh NAME -DFOO=1 ...
# -> tools/_build/NAME/liberyn_b200.so (use with LD_LIBRARY_PATH=tools/_build/NAME tools/_build/microbench_prod ...)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
OUT=tools/_build/$NAME
mkdir -p $OUT
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $@"
for f in abi_core k_swap k_swap_split k_shard k_rj host_job k_stage k_mt; do
  nvcc $FLAGS -c eryn_b200/csrc/$f.cu -o $OUT/$f.o &
done
for k in 0 1 2; do
  nvcc $FLAGS -DEB_ONLY_LIKE=$k -c eryn_b200/csrc/k_stretch.cu -o $OUT/k_stretch_$k.o &
  nvcc $FLAGS -DEB_ONLY_LIKE=$k -c eryn_b200/csrc/k_gauss.cu -o $OUT/k_gauss_$k.o &
done
wait
nvcc -shared -o $OUT/liberyn_b200.so $OUT/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built $OUT/liberyn_b200.so
