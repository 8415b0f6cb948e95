# round-2 profiling pass on one GPU: launch lists, ncu --set full of the C2 kernels and of the config-5 kernels, sanitizer
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stretch_step_kernel|pt_swap_kernel" -s 9 -c 6 -o gpurun_out/r02_ncu_full_c2 python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_full_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mb_group_stretch_kernel|mb_rj_kernel" -s 4 -c 4 -o gpurun_out/r02_ncu_full_c5 python tools/bench_c5.py --iters 4 --cpu-iters 0 > gpurun_out/ncu_full_c5.log 2>&1
cat > /tmp/api_run.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from tests.test_gpu_api import make_sampler, stretch_only
smp, _ = make_sampler(16, 4096, 8, stretch_only)
x0 = np.random.RandomState(3).uniform(-3, 3, size=(16, 4096, 8))
smp.run_mcmc(x0, 2, thin_by=10)
smp.run_mcmc(None, 3, thin_by=10)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_api.csv python /tmp/api_run.py > gpurun_out/ncu_api.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x -k "tight or mt_d4 or gibbs_d5 or wraparound or gauss_modes_d8" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_memcheck.log
python tools/bench_c5.py --iters 40 --cpu-iters 0 | tee gpurun_out/r02_bench_c5.txt
