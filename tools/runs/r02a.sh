mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02a_topo.txt 2>&1
(cat /proc/self/status | grep -i allowed; ls /sys/devices/system/node/; cat /sys/devices/system/node/node*/cpulist; nproc; free -g; lscpu | head -30) > gpurun_out/r02a_host.txt 2>&1
tools/h2d_probe --gpus 1 --mb 4096 --modes default,local,node0,node1,interleave --streams 1,2 > gpurun_out/r02a_h2d.json 2> gpurun_out/r02a_h2d.err
python tools/timeline_check.py --frames 5120 --streams 4 > gpurun_out/r02a_timeline.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
