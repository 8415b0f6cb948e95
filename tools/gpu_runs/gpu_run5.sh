mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for v in "EB_NO_PDL=0" "EB_NO_PDL=1"; do
for shape in "16 4096 8" "2 32768 8" "32 16384 20"; do
  echo "== $v $shape"
  env $v timeout 120 tools/_build/microbench $shape 2>&1 | grep -E "eb_stretch_step \(both|eb_gaussian|iteration|stretch CTA"
done; done
