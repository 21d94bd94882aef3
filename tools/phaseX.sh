timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_frontend_gpu.py -q > gpurun_out/pytestX0.log 2>&1; echo fetests=$?; tail -2 gpurun_out/pytestX0.log
timeout 300 python bench.py --workload frontend --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_Xf.json 2> gpurun_out/bench_X.err; echo benchf=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_Xf.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])"
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_X.json 2>> gpurun_out/bench_X.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_X.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
