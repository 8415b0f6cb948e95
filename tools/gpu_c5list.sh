mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c5.csv python tools/bench_c5.py --iters 12 --cpu-iters 0 > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log
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
for k,v in agg.items(): print(f"{k:72s} n={len(v):4d} avg={sum(v)/len(v)/1000:9.2f} us total={sum(v)/1e6:8.3f} ms")
PY
