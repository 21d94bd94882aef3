timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestR.log 2>&1; echo pytest=$?; grep -E "passed|failed|^FAILED" gpurun_out/pytestR.log | tail
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_R.json 2> gpurun_out/bench_R.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_R.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
