mkdir -p gpurun_out
timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-extra --cpu-sample 0 --opt subbatch=1000 > gpurun_out/r03r_bench_sb1000.json 2> gpurun_out/r03r_bench_sb1000.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-extra --cpu-sample 0 --opt subbatch=2000 > gpurun_out/r03r_bench_sb2000.json 2>> gpurun_out/r03r_bench_sb1000.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-extra --cpu-sample 0 > gpurun_out/r03r_plain.json 2> gpurun_out/r03r_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/r03r_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --cpu-sample 0 > gpurun_out/r03r_ncu.log 2>&1
