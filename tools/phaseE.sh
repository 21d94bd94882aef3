set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestE.log 2>&1; echo pytest=$?
grep -E "passed|failed|^FAILED" gpurun_out/pytestE.log | tail -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smokeE.log 2>&1; echo smoke=$?; tail -2 gpurun_out/smokeE.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_E.json 2> gpurun_out/bench_E.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_E.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d.get('ms_by_entry_point_eager'), d['gpu_launches_per_step'])"
timeout 300 python tools/gpu_bench_layers.py 160 > gpurun_out/layers_E.log 2>&1; tail -9 gpurun_out/layers_E.log; grep -E "enc1|dec8" gpurun_out/layers_E.log
