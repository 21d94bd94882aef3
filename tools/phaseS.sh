timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestS.log 2>&1; echo pytest=$?; grep -E "passed|failed|^FAILED" gpurun_out/pytestS.log | tail
timeout 300 python bench.py --workload inference --steps 20 --warmup 5 > gpurun_out/bench_inference_r01.json 2> gpurun_out/bench_S.err; echo benchi=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_inference_r01.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'])"
