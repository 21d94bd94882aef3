set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestH.log 2>&1; echo pytest=$?
grep -E "passed|failed|^FAILED" gpurun_out/pytestH.log | tail -30
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_H.json 2> gpurun_out/bench_H.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_H.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['gpu_launches_per_step'])"
timeout 300 python tools/gpu_bench_layers.py 160 > gpurun_out/layers_H.log 2>&1; head -22 gpurun_out/layers_H.log; tail -1 gpurun_out/layers_H.log
