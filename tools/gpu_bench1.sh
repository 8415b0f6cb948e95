mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_integration_stub.py -m gpu -q -x -k "run_host or integration or error_behaviour" 2>&1 | tail -5
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -2
