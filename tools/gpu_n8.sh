mkdir -p gpurun_out
bash tools/gpu_breakdown8.sh 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r02_bench_n8_auto.json 2> gpurun_out/r02_bench_n8_auto.err; echo "n8 rc=$?"
