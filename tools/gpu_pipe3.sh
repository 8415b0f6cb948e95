mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host_pipeline.py tests/test_gpu_parity.py::test_run_host_matches_oracle -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_pipe_pytest.log
EB_PROBE_N=100 timeout 300 python tools/e2e_probe.py 2>&1 | tee gpurun_out/r02_e2e_probe.txt
echo "--- no PDL between uploads"
EB_HOST_PDL=0 EB_PROBE_N=100 EB_PROBE_ONLY="wave G=4 graph;wave G=8 graph;wave G=16 graph" timeout 300 python tools/e2e_probe.py 2>&1 | tee -a gpurun_out/r02_e2e_probe.txt
echo "--- 16 / 64 copy CTAs"
EB_HOST_COPY_CTAS=16 EB_PROBE_N=100 EB_PROBE_ONLY="wave G=4 graph;wave G=8 graph" timeout 300 python tools/e2e_probe.py 2>&1 | tee -a gpurun_out/r02_e2e_probe.txt
EB_HOST_COPY_CTAS=64 EB_PROBE_N=100 EB_PROBE_ONLY="wave G=4 graph;wave G=8 graph" timeout 300 python tools/e2e_probe.py 2>&1 | tee -a gpurun_out/r02_e2e_probe.txt
EB_PROBE_STAMPS=1 EB_PROBE_ONLY="wave G=8 graph" timeout 300 python tools/e2e_probe.py 2>&1 | tail -2 | tee gpurun_out/r02_e2e_probe_stamps.txt
