mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
