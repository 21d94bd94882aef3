timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestY.log 2>&1; echo pytest=$?; grep -E "passed|failed|^FAILED" gpurun_out/pytestY.log | tail
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_Y.json 2> gpurun_out/bench_Y.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_Y.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_launches_per_step'])"
