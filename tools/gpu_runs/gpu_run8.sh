mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
for shape in "16 4096 8" "8 8192 8" "32 16384 20"; do
  echo "== $shape"
  timeout 120 tools/_build/microbench $shape 2>&1 | grep -E "eb_stretch_step \(both|eb_pt_swap|iteration|swap    CTA"
done
