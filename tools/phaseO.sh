timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestO.log 2>&1; echo pytest=$?; grep -E "passed|failed|^FAILED" gpurun_out/pytestO.log | tail
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_O.json 2> gpurun_out/bench_O.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_O.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_launches_per_step'])"
timeout 300 python tools/gpu_bench_layers.py 160 2>&1 | grep -E "^enc|^dec|totals"
