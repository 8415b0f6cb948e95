mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
which compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_rj.py -m gpu -x -q -k "c1 or tight or odd or T40 or W_small or c3_small or pt_kat2 or small or backend" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_rj.py -m gpu -x -q -k "c1 or odd or T24 or W_small or c5_small or backend" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/sanitizer_racecheck.log
