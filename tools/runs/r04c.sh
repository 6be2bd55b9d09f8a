mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r04c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r04c_pytest.log
python tools/profile_run.py --frames 2048 --iters 3 --streams 1 --stages 2 2>&1 | grep -o "ms_screen [0-9.]*\|checksum [0-9]*"
LM_WHATIF_SKIP=123 LM_WHATIF_S2=128 timeout 60 python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 2>&1 | grep "k_screen2 pair" | tail -13 | awk "{s+=\$9; n+=\$12; c++} END {printf \"  avg loop cycles %.0f, units %.1f, per MMA %.2f\n\", s/c, n/c, s/n/60}"
python bench.py --steps 4 --warmup 3 --no-e2e --no-extra --cpu-sample 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['roofline']['frac'], d['roofline']['launch_ms'])"
