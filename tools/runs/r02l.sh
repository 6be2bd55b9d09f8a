mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
for s in 512 1024; do echo "LM_NMS_SMALL=$s"; LM_NMS_SMALL=$s python tools/whatif.py 10000 0,32 4; done > gpurun_out/r02l_whatif.txt 2>&1
python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02l_plain.log 2>&1 &&
ncu -k regex:^k_nms --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/r02l_kernels.csv python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02l_ncu.log 2>&1
