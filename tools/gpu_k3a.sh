mkdir -p gpurun_out
for i in 1 2; do tools/_build/microbench 16 4096 8 2>&1 | grep "swap marks"; done | tee gpurun_out/r02_micro_c2_marks3.txt
tools/_build/microbench_prod 16 4096 8 2>&1 | grep -i "stretch_step (both\|eb_pt_swap  \|iteration\|sharded swap" | tee gpurun_out/r02_micro_c2_k3ticket.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_pipeline.py -x -q -m gpu 2>&1 | tail -4
