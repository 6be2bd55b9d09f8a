mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_screen.py tests/test_gpu_pass1_base.py -m gpu -q -x > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
for rc in 1536 768 640 512; do echo "LM_TAIL_RUNCAP=$rc"; LM_TAIL_RUNCAP=$rc python tools/whatif.py 10000 0 4; done > gpurun_out/r02r_whatif.txt 2>&1
LM_TAIL_RUNCAP=640 python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02r_plain.log 2>&1 &&
LM_TAIL_RUNCAP=640 ncu -k regex:^k_tail --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem --clock-control none -s 4 -c 4 --csv --log-file gpurun_out/r02r_kernels.csv python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02r_ncu.log 2>&1
