mkdir -p gpurun_out
N=${1:-8}
export EB_BREAKDOWN_MODES=split
for shape in c2 c4; do
EB_BREAKDOWN_SHAPE=$shape ERYN_B200_LIB=$PWD/tools/_build/liberyn_b200_prof.so timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/shard_breakdown.py 2>&1 | grep "^\[" | grep -E "rank 0|rank 3|rank 7" | tee gpurun_out/r02_breakdown_n${N}_$shape.txt
done
