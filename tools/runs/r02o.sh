mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_cpp.py -q -x > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
