mkdir -p gpurun_out
for SHAPE in "16 4096 8 0" "16 4096 8 1" "64 4096 8 0" "4 16384 20 2"; do
  for ol in 0 1; do
    echo "== shape $SHAPE EB_K1_ONE_LAUNCH=$ol"
    EB_K1_ONE_LAUNCH=$ol timeout 120 tools/_build/microbench_prod $SHAPE 2>&1 | grep -E "eb_stretch_step \(both|eb_pt_swap  |iteration"
  done
done 2>&1 | tee gpurun_out/r02_k1_one_launch.txt
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_n1.err
