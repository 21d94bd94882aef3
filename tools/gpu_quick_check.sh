# Quick GPU regression: all GPU tests + a 20-step train bench (no CPU baseline).
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_quick.log 2>&1; echo pytest=$?; grep -E "passed|failed|^FAILED" gpurun_out/pytest_quick.log | tail
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_launches_per_step'])"
