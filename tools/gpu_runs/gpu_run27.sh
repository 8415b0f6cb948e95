mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_c5.py --nt 500 --iters 50 --cpu-iters 0 2>&1 | tail -2 | tee gpurun_out/bench_c5.txt
timeout 120 tools/_build/microbench_prod 32 16384 20 2 2>&1 | grep -E "eb_stretch_step|eb_pt_swap|eb_gauss|iteration" | tee gpurun_out/micro_c4.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stretch_step_kernel" -s 6 -c 2 -f -o gpurun_out/r01_ncu_full_c4_k1 tools/_build/microbench_prod 32 16384 20 2 > gpurun_out/ncu_c4.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
