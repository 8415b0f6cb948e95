mkdir -p gpurun_out
EB_HOST_UPLOAD=1 timeout 600 python -m pytest tests/test_gpu_host_pipeline.py -x -q -m gpu 2>&1 | tail -3
echo "--- upload mode 1 (copy engine + arrival polling), 32 poll CTAs"
EB_HOST_UPLOAD=1 EB_PROBE_N=100 EB_PROBE_ONLY="wave G=2 graph;wave G=4 graph;wave G=6 graph;wave G=8 graph;wave G=16 graph;wave 2,6,6,2;wave 1,3,4,4,3,1" timeout 300 python tools/e2e_probe.py 2>&1 | tee gpurun_out/r02_e2e_probe_dma.txt
echo "--- 8 poll CTAs"
EB_HOST_POLL_CTAS=8 EB_HOST_UPLOAD=1 EB_PROBE_N=100 EB_PROBE_ONLY="wave G=4 graph;wave G=6 graph;wave G=8 graph" timeout 300 python tools/e2e_probe.py 2>&1 | tee -a gpurun_out/r02_e2e_probe_dma.txt
EB_HOST_UPLOAD=1 EB_PROBE_STAMPS=1 EB_PROBE_ONLY="wave G=6 graph" timeout 300 python tools/e2e_probe.py 2>&1 | tail -2 | cut -c1-1800 | tee -a gpurun_out/r02_e2e_probe_dma.txt
