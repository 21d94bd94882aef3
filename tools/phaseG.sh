set -x
timeout 300 python -m pytest tests/test_flow_fused_gpu.py -q -x > gpurun_out/pytestG0.log 2>&1; echo flowtests=$?; tail -3 gpurun_out/pytestG0.log
timeout 300 python tools/gpu_bench_small.py 160 > gpurun_out/small_G.log 2>&1; cat gpurun_out/small_G.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestG.log 2>&1; echo pytest=$?
grep -E "passed|failed|^FAILED" gpurun_out/pytestG.log | tail -30
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_G.json 2> gpurun_out/bench_G.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_G.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('ms_by_entry_point_eager'), d['gpu_launches_per_step'])"
