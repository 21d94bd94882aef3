set -x
PGV_WGRAD_VARIANT=0 timeout 120 python tools/gpu_debug_wgrad.py > gpurun_out/wgrad_v0.log 2>&1; echo v0=$?
PGV_WGRAD_VARIANT=1 timeout 120 python tools/gpu_debug_wgrad.py > gpurun_out/wgrad_v1.log 2>&1; echo v1=$?
grep -E "variant" gpurun_out/wgrad_v0.log gpurun_out/wgrad_v1.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestC.log 2>&1; echo pytest=$?
grep -E "passed|failed|^FAILED" gpurun_out/pytestC.log | tail -30
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cl.json 2> gpurun_out/bench_cl.err; echo bench=$?
cut -c1-300 gpurun_out/bench_cl.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cl.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('ms_by_entry_point_eager'))"
