mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_rj.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for T in 32 64 128; do
echo "== T=$T"; timeout 100 tools/_build/microbench_prod $T 4096 8 2>&1 | grep -E "eb_pt_swap|sharded swap, world=1, no"
done | tee gpurun_out/micro_T.txt
timeout 100 tools/_build/microbench 128 4096 8 2>&1 | grep -E "swap marks|swap    CTA" | tee -a gpurun_out/micro_T.txt
