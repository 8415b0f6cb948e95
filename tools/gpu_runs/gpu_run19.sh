mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for shape in "16 4096 8" "32 16384 20"; do
  echo "== $shape"
  timeout 120 tools/_build/microbench_prod $shape 2>&1 | grep -E "eb_stretch_step \(both|eb_pt_swap|iteration"
done
