# K1 variants on one GPU: parity of the lane-split kernel, then timings (tools/microbench.cu on the production library)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lane_split or config4" 2>&1 | tail -15 > gpurun_out/r02_k1_lanes_pytest.log
cat gpurun_out/r02_k1_lanes_pytest.log
for shape in "32 16384 20 2" "16 4096 8 0" "16 4096 8 1" "16 16384 8 0" "64 4096 8 0"; do
  for lpw in 1 2 4; do
    echo "== shape $shape EB_K1_LPW=$lpw"
    EB_K1_LPW=$lpw timeout 120 tools/_build/microbench_prod $shape 2>&1 | grep -E "eb_stretch_step \(both|iteration|eb_pt_swap  "
  done
done 2>&1 | tee gpurun_out/r02_k1_lanes_micro.txt
