timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_conv_cl_gpu.py -q -k "tensor_core_convs or big_linear or block_forward" > gpurun_out/pytestW0.log 2>&1; echo convtests=$?; tail -2 gpurun_out/pytestW0.log
timeout 300 python tools/gpu_bench_layers.py 160 2>&1 | grep -E "^enc|^dec|totals"
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_W.json 2> gpurun_out/bench_W.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_W.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'])"
