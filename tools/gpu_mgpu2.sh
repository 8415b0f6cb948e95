mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r02_bench_n2_auto.json 2> gpurun_out/r02_bench_n2_auto.err; echo "n2 rc=$?"; tail -1 gpurun_out/r02_bench_n2_auto.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n2_auto.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d.get("parity"), d["extra"].get("c4_strong",{}).get("speedup_vs_1gpu"))
PY
