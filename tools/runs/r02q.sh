mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q_pytest.log
python tools/whatif.py 10000 0,2,123 4 > gpurun_out/r02q_whatif.txt 2>&1
LM_WHATIF_DETCAP=1024 LM_TAIL_RUNCAP=128 python tools/whatif.py 10000 0,16,32 4 > gpurun_out/r02q_whatif_cores.txt 2>&1
python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02q_plain.log 2>&1 &&
ncu -k regex:^k_ --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/r02q_kernels.csv python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02q_ncu.log 2>&1
