mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"
tail -c 600 gpurun_out/bench_n1.err
for comm in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --comm $comm > gpurun_out/bench_n2_$comm.json 2> gpurun_out/bench_n2_$comm.err; echo "bench2 $comm rc=$?"
tail -c 1500 gpurun_out/bench_n2_$comm.err
done
cat gpurun_out/bench_n1.json gpurun_out/bench_n2_p2p.json gpurun_out/bench_n2_nccl.json | cut -c1-1500
