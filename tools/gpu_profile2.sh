# round-2 profiling, second pass: launch list of the wavefront eb_run_host call, ncu --set full of the resident kernel
mkdir -p gpurun_out
EB_PROBE_ONLY="wave default" EB_PROBE_N=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_e2e.csv python tools/e2e_probe.py > gpurun_out/ncu_e2e.log 2>&1
tail -2 gpurun_out/ncu_e2e.log
cat > /tmp/res_run.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from eryn_b200.device import DeviceContext
from eryn_b200.likelihood import GaussianLikelihood
from eryn_b200.prior import ProbDistContainer, uniform_dist
from eryn_b200.state import State
T, W, d = 16, 4096, 8
r = np.random.RandomState(0)
A = r.randn(d, d)
ctx = DeviceContext(ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)}), GaussianLikelihood(np.zeros(d), np.linalg.inv(A @ A.T / d + np.eye(d))), rng="philox", seed=3)
ds = ctx.upload(State({"model_0": r.uniform(-3, 3, size=(T, W, 1, d))}), betas=torch.from_numpy(np.geomspace(1.0, 1e-3, T)).to(ctx.device))
ctx.eval_state(ds)
ad = dict(adaptive=True, stop_adaptation=-1, adaptation_lag=10000.0, adaptation_time=100.0)
for _ in range(3):
    ctx.resident_run(ds, 2.0, 10, adapt=ad)
torch.cuda.synchronize()
ctx.check_error()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_resident python /tmp/res_run.py > gpurun_out/ncu_resident.log 2>&1
tail -3 gpurun_out/ncu_resident.log
ls -la gpurun_out/r02_ncu_resident.ncu-rep
