mkdir -p gpurun_out
for w in 32; do echo "== LM_WHATIF_S2=$w (other stages skipped)"; LM_WHATIF_SKIP=123 LM_WHATIF_S2=$w timeout 60 python tools/profile_run.py --frames 2048 --iters 3 --streams 1 --stages 2 --subbatch 2048 2>&1 | grep -o "ms_screen [0-9.]*\|checksum [0-9]*"; done > gpurun_out/r02y_s2_whatif5.txt 2>&1
