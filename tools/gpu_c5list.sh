mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c5.csv python tools/bench_c5.py --iters 12 --cpu-iters 0 > gpurun_out/ncu_c5.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_launches_c5.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            k=d['Kernel Name'][:70]
            v=float(d['Metric Value'].replace(',',''))
            agg.setdefault(k,[]).append(v)
for k,v in agg.items():
    if k.startswith("eb::") or "pt_swap" in k: print(f"{k:72s} n={len(v):4d} avg={sum(v)/len(v)/1000:9.2f} us")
PY
tools/_build/microbench_prod 64 4096 8 2>&1 | grep -i "eb_pt_swap  \|iteration" | tee gpurun_out/r02_micro_T64.txt
tools/_build/microbench_prod 128 2048 8 2>&1 | grep -i "eb_pt_swap  \|iteration" | tee -a gpurun_out/r02_micro_T64.txt
python - <<'PY'
import cProfile, pstats, sys, io
sys.path.insert(0, ".")
sys.argv = ["bench_c5.py", "--iters", "30", "--cpu-iters", "0"]
import runpy
pr = cProfile.Profile()
pr.enable()
try:
    runpy.run_path("tools/bench_c5.py", run_name="__main__")
except SystemExit:
    pass
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:5000])
PY
