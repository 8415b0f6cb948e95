mkdir -p gpurun_out
export EB_BREAKDOWN_MODES=fused
ERYN_B200_LIB=$PWD/tools/_build/liberyn_b200_prof.so timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/shard_breakdown.py 2>&1 | grep "^\[" | tee gpurun_out/breakdown_n2.txt
