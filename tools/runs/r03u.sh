mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pass1.py tests/test_gpu_pass1_tm.py tests/test_gpu_pass1_base.py tests/test_host_cpp.py -m gpu -q -x > gpurun_out/r03u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03u_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-extra --cpu-sample 0 > gpurun_out/r03u_bench.json 2> gpurun_out/r03u_bench.err
