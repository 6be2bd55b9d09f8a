mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pass1.py tests/test_gpu_pass1_tm.py -m gpu -q -x > gpurun_out/r03v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03v_pytest.log
timeout 200 python tools/pass1_check.py > gpurun_out/r03v_pass1.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 60 --csv --log-file gpurun_out/r03v_pass1_launches.csv python tools/pass1_check.py > gpurun_out/r03v_ncu.log 2>&1
