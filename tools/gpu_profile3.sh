mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stretch_step_kernel|pt_swap_kernel" -s 9 -c 6 -o gpurun_out/r02_ncu_full_c2_lazy python bench.py --steps 5 --warmup 3 --profile > gpurun_out/ncu_full_c2.log 2>&1
tail -2 gpurun_out/ncu_full_c2.log
