mkdir -p gpurun_out
python tools/profile_run.py --frames 2048 --iters 3 --streams 1 --stages 2 2>&1 | grep -o "ms_screen [0-9.]*\|checksum [0-9]*"
LM_WHATIF_SKIP=0 LM_WHATIF_S2=128 timeout 60 python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 2>&1 | grep "k_screen2 pair" | tail -13 | sed -n "1p;9p"
python tools/whatif.py 10000 0,123,122 4
