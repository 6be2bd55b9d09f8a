mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "rc=$?" >> gpurun_out/r02m_bench.err
