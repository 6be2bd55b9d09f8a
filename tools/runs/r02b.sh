mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
