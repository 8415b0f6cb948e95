timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c4 or d30 or T40_d20" 2>&1 | tail -2
timeout 120 tools/_build/microbench_prod 32 16384 20 2 2>&1 | grep -E "shape|eb_stretch_step|eb_gaussian|eb_pt_swap|eb_eval|iteration"
