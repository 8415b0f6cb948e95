mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for shape in "16 4096 8" "8 8192 8" "2 32768 8" "32 16384 20"; do
  timeout 120 tools/_build/microbench $shape > "gpurun_out/micro_$(echo $shape | tr ' ' '_').txt" 2>&1
  grep -v "^empty\|cluster8" "gpurun_out/micro_$(echo $shape | tr ' ' '_').txt"
done
