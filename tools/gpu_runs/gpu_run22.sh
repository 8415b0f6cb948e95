mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 120 tools/_build/microbench_prod 16 4096 8 2>&1 | grep -E "eb_stretch_step \(both|eb_pt_swap|iteration"
