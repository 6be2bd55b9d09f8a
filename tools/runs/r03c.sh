mkdir -p gpurun_out
timeout 300 python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages 2 2>&1 | grep -o "ms_screen [0-9.]*\|checksum [0-9]*" > gpurun_out/r03c_plain.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r03c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03c_pytest.log
timeout 200 python tools/whatif.py 10000 0,123,11,112 4,8 > gpurun_out/r03c_whatif.txt 2>&1
