timeout 300 python tools/gpu_bench_layers.py 160 2>&1 | grep -E "^enc|^dec|totals"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestL.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytestL.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_L.json 2> gpurun_out/bench_L.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_L.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"
