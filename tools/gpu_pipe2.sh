mkdir -p gpurun_out
EB_PROBE_STAMPS=1 EB_PROBE_ONLY="wave G=1 graph;wave G=2 graph;wave G=4 graph;wave G=8 graph" timeout 300 python tools/e2e_probe.py 2>&1 | tee gpurun_out/r02_e2e_probe_stamps.txt
