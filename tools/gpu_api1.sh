python - <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
wl = bench.workload("c2")
dev = torch.device("cuda", 0)
for kw in (dict(store=True), dict(nsteps=10, thin_by=100, store=True), dict(store=False), dict(store=True)):
    r = bench.bench_api(wl, dev, **kw)
    print({k: r[k] for k in ("value", "ms_per_step", "thin_by", "store", "stored_steps")}, flush=True)
PY
cat /sys/kernel/mm/transparent_hugepage/enabled; nproc; uname -r
