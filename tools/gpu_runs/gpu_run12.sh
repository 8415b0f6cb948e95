mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
for shape in "16 4096 8" "8 8192 8" "32 16384 20"; do
  echo "== $shape"
  timeout 120 tools/_build/microbench_prod $shape 2>&1 | grep -E "eb_stretch_step \(both|eb_pt_swap|eb_gauss|iteration"
done
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; tail -c 400 gpurun_out/bench_n1.err
cut -c1-1800 gpurun_out/bench_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
tail -c 600 gpurun_out/bench_n2.err; cut -c1-400 gpurun_out/bench_n2.json
