mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu.log | head -30
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pt_kat2" > gpurun_out/sanitizer_kat2.log 2>&1; echo "sanitizer rc=$?"
grep -E "Invalid|at .*\.cu|by thread|Address|ERROR SUMMARY" gpurun_out/sanitizer_kat2.log | head -30
