mkdir -p gpurun_out
for n in 8 4; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"
tail -c 300 gpurun_out/bench_n$n.err | grep -v "^\*\|OMP_NUM"; cut -c1-330 gpurun_out/bench_n$n.json
done
