mkdir -p gpurun_out
run() { # N comm
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1 bench.py --gpus $1 --steps 200 --warmup 10 --comm $2 > gpurun_out/r02_bench_n$1_$2.json 2> gpurun_out/r02_bench_n$1_$2.err; echo "n$1 $2 rc=$?"; tail -c 400 gpurun_out/r02_bench_n$1_$2.err | tail -3
}
run 8 split
run 8 fused
run 4 split
