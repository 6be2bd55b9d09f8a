mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra --cpu-sample 0 > gpurun_out/r02t_bench_n2.json 2> gpurun_out/r02t_bench_n2.err
