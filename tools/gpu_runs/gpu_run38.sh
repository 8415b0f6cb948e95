mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "combine or kat or c2_full or T128 or T70 or T8 or c3_small" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest parity rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python -m pytest tests/test_mgpu.py -m gpu -x -q -k "fused and (4-256 or 5-99 or 16-4096 or 128-48)" > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$?"; tail -3 gpurun_out/pytest_mgpu.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
tail -c 200 gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.json
