# streaming stretch kernel: parity, timings of all variants at config-4 size, ncu of the best; then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lane_split or config4" 2>&1 | tail -25 > gpurun_out/r02_k1_stream_pytest.log
cat gpurun_out/r02_k1_stream_pytest.log
run() { echo "== $1"; shift; env "$@" timeout 120 tools/_build/microbench_prod $SHAPE 2>&1 | grep -E "eb_stretch_step \(both|eb_stretch_step, no count"; }
for SHAPE in "32 16384 20 2" "16 16384 8 2"; do
  echo "#### shape $SHAPE"
  run "thread-per-walker" EB_K1_LPW=1
  run "lanes4" EB_K1_LPW=4
  run "stream minb3" EB_K1_LPW=8
  run "stream minb4" EB_K1_LPW=8 LD_LIBRARY_PATH=tools/_build/s4
  run "stream minb2" EB_K1_LPW=8 LD_LIBRARY_PATH=tools/_build/s2
done 2>&1 | tee gpurun_out/r02_k1_stream_variants.txt
SHAPE="32 16384 20 2"
EB_K1_LPW=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_stream -s 4 -c 2 -o gpurun_out/r02_ncu_c4_stream tools/_build/microbench_prod $SHAPE > gpurun_out/ncu_stream.log 2>&1
EB_K1_LPW=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_lanes -s 4 -c 2 -o gpurun_out/r02_ncu_c4_lanes tools/_build/microbench_prod $SHAPE > gpurun_out/ncu_lanes.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_pytest_gpu.log; tail -30 gpurun_out/r02_pytest_gpu.log
