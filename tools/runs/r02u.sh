mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest.log
python tools/whatif.py 10000 0 4,6,8 > gpurun_out/r02u_whatif.txt 2>&1
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "rc=$?" >> gpurun_out/r02u_bench.err
python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 > gpurun_out/r02u_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:k_screen2 -s 2 -c 1 -o gpurun_out/r02u_screen2 python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 > gpurun_out/r02u_ncu.log 2>&1
