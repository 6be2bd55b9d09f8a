mkdir -p gpurun_out
for s in 1024 512 256; do echo "LM_NMS_SMALL=$s"; LM_NMS_SMALL=$s python tools/whatif.py 10000 0,32 4; done > gpurun_out/r02j_nms_small.txt 2>&1
