mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pass1_tm.py tests/test_gpu_pass1.py tests/test_gpu_pass1_base.py -q -x > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
