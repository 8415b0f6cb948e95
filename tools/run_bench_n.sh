N=$1
for comm in split fused; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --comm $comm > gpurun_out/r02_bench_n${N}_${comm}.json 2> gpurun_out/r02_bench_n${N}_${comm}.err
tail -c 600 gpurun_out/r02_bench_n${N}_${comm}.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n${N}_${comm}.json"))
    print("${comm}", d["n_gpus"], d["value"], d["ms_per_step"], d["extra"]["ms_per_step_resident_no_flush"])
except Exception as e: print("fail", e)
PY
done
