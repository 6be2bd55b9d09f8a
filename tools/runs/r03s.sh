mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r03s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03s_pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r03s_bench.json 2> gpurun_out/r03s_bench.err; echo "rc=$?" >> gpurun_out/r03s_bench.err
