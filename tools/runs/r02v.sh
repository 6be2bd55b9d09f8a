mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pass1_tm.py tests/test_host_cpp.py -m gpu -q -x > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v_pytest.log
timeout 600 python tools/pass1_tm_check.py 2048 > gpurun_out/r02v_pass1_tm.txt 2>&1
