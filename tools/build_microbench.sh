#!/bin/bash
# Profiling build of the library (phase timers on) + the microbenchmark binary.  Outputs are git-ignored.
set -e
cd "$(dirname "$0")/.."
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -DEB_PHASE_TIMERS"
mkdir -p tools/_build
for f in abi_core k_stretch k_gauss k_swap k_swap_split k_shard k_rj host_job k_stage k_mt; do
  nvcc $FLAGS -c eryn_b200/csrc/$f.cu -o tools/_build/$f.o &
done
wait
nvcc -shared -o tools/_build/liberyn_b200_prof.so tools/_build/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include tools/microbench.cu -L tools/_build -leryn_b200_prof \
     -Xlinker -rpath -Xlinker '$ORIGIN' -o tools/_build/microbench
echo built tools/_build/microbench
# the same benchmark against the PRODUCTION library (no phase timers): these are the timings to quote
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -DNO_MARKS tools/microbench.cu -L eryn_b200/lib -leryn_b200 \
     -Xlinker -rpath -Xlinker '$ORIGIN/../../eryn_b200/lib' -o tools/_build/microbench_prod
echo built tools/_build/microbench_prod
