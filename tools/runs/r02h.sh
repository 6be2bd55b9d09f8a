mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02h_plain.log 2>&1 &&
ncu -k regex:^k_ --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/r02h_kernels.csv python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 > gpurun_out/r02h_ncu.log 2>&1
