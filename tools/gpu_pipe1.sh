mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host_pipeline.py tests/test_gpu_parity.py::test_run_host_matches_oracle -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_pipe_pytest.log
timeout 300 python tools/e2e_probe.py 2>&1 | tee gpurun_out/r02_e2e_probe.txt
