mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nms_reference.py tests/test_gpu_screen.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r03k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03k_pytest.log
python tools/whatif.py 10000 0,91,75,112 4,8 > gpurun_out/r03k_whatif.txt 2>&1
LM_TAIL_RUNCAP=640 python tools/whatif.py 10000 0 4,8 >> gpurun_out/r03k_whatif.txt 2>&1
