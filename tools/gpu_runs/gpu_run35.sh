mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_rj.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; tail -c 300 gpurun_out/bench_n1.err
cut -c1-300 gpurun_out/bench_n1.json
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 100 tools/_build/microbench_prod 16 4096 8 2>&1 | grep -E "eb_stretch_step \(both|eb_pt_swap|iteration|sharded swap, world=1, no" | tee gpurun_out/micro_c2.txt
timeout 100 tools/_build/microbench_prod 128 4096 8 2>&1 | grep -E "eb_pt_swap|sharded swap, world=1, no" | sed 's/^/T128 /' | tee -a gpurun_out/micro_c2.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
