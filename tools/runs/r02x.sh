mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_pytest.log
for st in 2 3 4; do echo "== stages $st"; python tools/profile_run.py --frames 1024 --iters 3 --streams 1 --stages $st 2>&1 | tail -4; done > gpurun_out/r02x_plain.log 2>&1
python tools/whatif.py 10000 0,123 4 > gpurun_out/r02x_whatif.txt 2>&1
