mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r03q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03q_pytest.log
