mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r03l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03l_pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r03l_bench.json 2> gpurun_out/r03l_bench.err; echo "rc=$?" >> gpurun_out/r03l_bench.err
python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 > gpurun_out/r03l_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:k_screen2 -s 2 -c 1 -o gpurun_out/r03l_screen2 python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 > gpurun_out/r03l_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 28 -c 56 --csv --log-file gpurun_out/r03l_kernels.csv python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 > gpurun_out/r03l_ncu2.log 2>&1
