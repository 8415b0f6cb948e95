mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_rj.py tests/test_gpu_parity.py tests/test_gpu_host_pipeline.py -x -q -m gpu 2>&1 | tail -5
python tools/bench_c5.py --iters 40 --cpu-iters 0 2>&1 | tail -2 | tee gpurun_out/r02_bench_c5.txt
tools/_build/microbench_prod 64 4096 8 2>&1 | grep -i "eb_pt_swap  \|iteration" | tee gpurun_out/r02_micro_T64.txt
tools/_build/microbench_prod 128 2048 8 2>&1 | grep -i "eb_pt_swap  \|iteration" | tee -a gpurun_out/r02_micro_T64.txt
