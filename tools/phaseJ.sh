for v in 4 0; do echo "== variant $v"; PGV_CONV_VARIANT=$v timeout 300 python tools/gpu_bench_layers.py 160 2>&1 | grep -E "^enc|^dec|totals"; done
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_conv_cl_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py -q > gpurun_out/pytestJ.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytestJ.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k tensor_core_convs > gpurun_out/pytestJ$i.log 2>&1; echo rep$i=$?; done
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_J.json 2> gpurun_out/bench_J.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_J.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"
