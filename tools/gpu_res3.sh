mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_api.py tests/test_gpu_host_pipeline.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02_resident_pytest.log
