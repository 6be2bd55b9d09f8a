mkdir -p gpurun_out
python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02d_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r02d_kernels.csv python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02d_ncu.log 2>&1
