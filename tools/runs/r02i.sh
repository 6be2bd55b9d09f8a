mkdir -p gpurun_out
python tools/whatif.py 10000 > gpurun_out/r02i_whatif.txt 2>&1
