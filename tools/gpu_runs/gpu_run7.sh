mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
for v in 0 1 2 3 4 5 6 7; do
  echo "== EB_PDL_MASK=$v"
  env EB_PDL_MASK=$v timeout 120 tools/_build/microbench 16 4096 8 2>&1 | grep -E "eb_stretch_step \(both|eb_pt_swap|iteration|swap    CTA"
done
