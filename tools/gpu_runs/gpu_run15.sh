mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_rj.py -m gpu -x -q > gpurun_out/pytest_rj.log 2>&1; echo "pytest rj rc=$?"
tail -40 gpurun_out/pytest_rj.log
