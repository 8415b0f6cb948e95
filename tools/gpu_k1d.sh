timeout 900 python -m pytest tests/test_gpu_lazy_adapt.py tests/test_gpu_api.py tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -4
