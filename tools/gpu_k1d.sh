tools/_build/microbench_prod 16 4096 8 2>&1 | grep -i "stretch_step (both\|eb_pt_swap  \|iteration (str"
tools/_build/microbench_prod 4 128 8 2>&1 | grep -i "stretch_step (both\|iteration (str"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lazy_adapt.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -3
