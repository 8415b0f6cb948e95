# 2 GPUs: parity of the sharded passes (all comm modes), then the in-kernel timeline of split vs fused
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_mgpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02_mgpu_n2.log; cat gpurun_out/r02_mgpu_n2.log
export EB_BREAKDOWN_MODES=split,fused
ERYN_B200_LIB=$PWD/tools/_build/liberyn_b200_prof.so timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/shard_breakdown.py 2>&1 | grep "^\[" | tee gpurun_out/r02_breakdown_n2.txt
