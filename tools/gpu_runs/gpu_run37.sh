mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mgpu.py -m gpu -x -q -k "fused" > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$?"; tail -3 gpurun_out/pytest_mgpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
tail -c 200 gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.json
export EB_BREAKDOWN_MODES=fused
ERYN_B200_LIB=$PWD/tools/_build/liberyn_b200_prof.so timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/shard_breakdown.py 2>&1 | grep "^\[" | tee gpurun_out/breakdown_n2.txt
