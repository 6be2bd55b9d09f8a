mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02p_topo.txt 2>&1
(nproc; cat /sys/devices/system/node/node*/cpulist; free -g | head -2) > gpurun_out/r02p_host.txt 2>&1
for g in 1 2 4 8; do tools/h2d_probe --gpus $g --mb 2048 --modes default --streams 1 --reps 3; done > gpurun_out/r02p_h2d.json 2> gpurun_out/r02p_h2d.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 4 --warmup 3 --no-extra --cpu-sample 0 > gpurun_out/r02p_bench_n8.json 2> gpurun_out/r02p_bench_n8.err
