mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -15
bash tools/gpu_api1.sh 2>&1 | tee gpurun_out/r02_api_prefault.txt
