mkdir -p gpurun_out
python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 > gpurun_out/r02g_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k 'regex:k_nms_bottom|k_corr_sparse|k_prep|k_tail$|k_nms_side|k_pair|k_minmax' -s 22 -c 11 -o gpurun_out/r02g_small python tools/profile_run.py --frames 1024 --iters 2 --streams 1 --stages 2 > gpurun_out/r02g_ncu.log 2>&1
ls -la gpurun_out/r02g_small.ncu-rep
