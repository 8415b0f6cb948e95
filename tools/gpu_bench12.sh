mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "n1 rc=$?"; tail -c 1500 gpurun_out/r02_bench_n1.err
for comm in split fused; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --comm $comm > gpurun_out/r02_bench_n2_$comm.json 2> gpurun_out/r02_bench_n2_$comm.err; echo "n2 $comm rc=$?"; tail -c 1500 gpurun_out/r02_bench_n2_$comm.err
done
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r02_bench_ref.err
