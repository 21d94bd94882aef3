timeout 300 python -m pytest tests/test_conv_cl_gpu.py -q -s -k big_linear > gpurun_out/pytestN0.log 2>&1; echo fctests=$?; grep -E "^fc |passed|failed|Error" gpurun_out/pytestN0.log | head
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestN.log 2>&1; echo pytest=$?; grep -E "passed|failed|^FAILED" gpurun_out/pytestN.log | tail
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_N.json 2> gpurun_out/bench_N.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_N.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_launches_per_step'])"
