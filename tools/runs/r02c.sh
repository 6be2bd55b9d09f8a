mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host_cpp or host_tracks or pass1 or cost" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
for o in "screen_stages=3" "screen_stages=4" "streams=6" "streams=8" "streams=8 --opt screen_stages=3"; do
  echo "== $o" >> gpurun_out/r02c_opts.txt
  timeout 200 python bench.py --steps 6 --warmup 3 --no-e2e --cpu-sample 0 --opt $o 2>>gpurun_out/r02c_opts.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['value'], d['ms_per_step'], r['launch_ms'], r['launch_ms_in_timed_region_overlapped'])" >> gpurun_out/r02c_opts.txt
done
