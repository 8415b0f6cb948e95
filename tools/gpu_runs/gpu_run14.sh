mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; tail -c 300 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stretch_step_kernel|pt_swap_kernel" -s 9 -c 6 -f -o gpurun_out/r01_ncu_full_c2_pdl python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
tail -c 300 gpurun_out/bench_n2.err
