mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; tail -c 300 gpurun_out/bench_n1.err
cut -c1-400 gpurun_out/bench_n1.json
timeout 120 tools/_build/microbench_prod 16 4096 8 2>&1 | grep -E "eb_stretch_step \(both|eb_pt_swap|iteration" | tee gpurun_out/micro_c2.txt
timeout 120 tools/_build/microbench_prod 128 4096 8 2>&1 | grep -E "eb_pt_swap|iteration" | sed 's/^/T128 /' | tee -a gpurun_out/micro_c2.txt
timeout 300 python tools/bench_c5.py --nt 500 --iters 50 --cpu-iters 0 2>&1 | tail -1 | tee gpurun_out/bench_c5.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stretch_step_kernel|pt_swap_kernel" -s 9 -c 6 -f -o gpurun_out/r01_ncu_full_c2_final python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
