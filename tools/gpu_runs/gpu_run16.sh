mkdir -p gpurun_out
timeout 600 python tools/bench_c5.py --nt 500 --iters 50 2>&1 | tail -3 | tee gpurun_out/bench_c5.txt
timeout 300 python bench.py --workload c3 --steps 100 --warmup 5 --profile 2>&1 | tail -1 | cut -c1-700
