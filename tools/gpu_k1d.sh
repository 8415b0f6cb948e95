timeout 900 python -m pytest tests/test_gpu_lazy_adapt.py tests/test_gpu_api.py tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -1 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
ex=d.pop("extra")
print(json.dumps({k:d[k] for k in ("value","ms_per_step","gpu_launches")}), json.dumps(d["e2e"]["value"]), json.dumps(d["roofline"]["frac"]), d["roofline"]["avg_launch_us"])
print("swap", json.dumps(ex.get("roofline_swap",{}).get("avg_launch_us")), "iter", json.dumps(ex.get("roofline_iteration",{}).get("frac")), "res_noflush", ex.get("ms_per_step_resident_no_flush"), "c4roof", ex.get("roofline_c4",{}).get("frac"))
print(json.dumps({k:ex[k].get("value") for k in ("api","api_thin100","api_not_stored","c3","c4","c5")}))
PY
