mkdir -p gpurun_out
(echo "detcap 1024"; LM_WHATIF_DETCAP=1024 python tools/whatif.py 10000 0,32 4; echo "detcap 2048"; LM_WHATIF_DETCAP=2048 python tools/whatif.py 10000 0 4; echo "default"; python tools/whatif.py 10000 0,32 4,6) > gpurun_out/r02s_whatif.txt 2>&1
python tools/timeline_check.py --frames 10240 --streams 4 > gpurun_out/r02s_timeline.txt 2>&1
