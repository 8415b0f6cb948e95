mkdir -p gpurun_out
tools/_build/microbench_prod 128 2048 8 2>&1 | grep -i "eb_pt_swap  \|iteration"
timeout 600 python -m pytest tests/test_gpu_parity_rj.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import cProfile, pstats, sys, io
sys.path.insert(0, ".")
sys.argv = ["bench_c5.py", "--iters", "60", "--cpu-iters", "0"]
import runpy
pr = cProfile.Profile()
pr.enable()
try:
    runpy.run_path("tools/bench_c5.py", run_name="__main__")
except SystemExit:
    pass
pr.disable()
s = io.StringIO()
st = pstats.Stats(pr, stream=s)
st.sort_stats("tottime").print_stats("eryn_b200|ctypes|numpy|torch/_tensor|_lib", 30)
print(s.getvalue()[-6000:])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats("eryn_b200", 25)
print(s.getvalue()[-5000:])
PY
