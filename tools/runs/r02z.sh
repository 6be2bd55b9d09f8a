mkdir -p gpurun_out
python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 > gpurun_out/r02z_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:k_screen2 -s 2 -c 1 -o gpurun_out/r02z_screen2 python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 > gpurun_out/r02z_ncu.log 2>&1
