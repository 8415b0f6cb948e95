# K1 variants on one GPU (production library): prefetch on/off, one thread per walker vs lane-split, round loop unrolled
mkdir -p gpurun_out
run() { # label, env...
  echo "== $1"; shift
  env "$@" timeout 120 tools/_build/microbench_prod $SHAPE 2>&1 | grep -E "eb_stretch_step \(both|eb_stretch_step, no count"
}
for SHAPE in "32 16384 20 2" "16 16384 8 0" "16 4096 8 0"; do
  echo "#### shape $SHAPE"
  run "lpw1 pf0" EB_K1_LPW=1 EB_K1_PREFETCH=0
  run "lpw1 pf1" EB_K1_LPW=1 EB_K1_PREFETCH=1
  run "lpw4 pf0" EB_K1_LPW=4 EB_K1_PREFETCH=0
  run "lpw4 pf1" EB_K1_LPW=4 EB_K1_PREFETCH=1
  run "lpw4 pf1 unroll2" EB_K1_LPW=4 EB_K1_PREFETCH=1 LD_LIBRARY_PATH=tools/_build/u2
  run "lpw4 pf0 unroll2" EB_K1_LPW=4 EB_K1_PREFETCH=0 LD_LIBRARY_PATH=tools/_build/u2
done 2>&1 | tee gpurun_out/r02_k1_variants.txt
SHAPE="32 16384 20 2"
EB_K1_LPW=4 EB_K1_PREFETCH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_lanes -s 4 -c 2 -o gpurun_out/r02_ncu_c4_lanes tools/_build/microbench_prod $SHAPE > gpurun_out/ncu_lanes.log 2>&1
EB_K1_LPW=1 EB_K1_PREFETCH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_step -s 4 -c 2 -o gpurun_out/r02_ncu_c4_thread tools/_build/microbench_prod $SHAPE > gpurun_out/ncu_thread.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
