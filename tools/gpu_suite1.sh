mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
