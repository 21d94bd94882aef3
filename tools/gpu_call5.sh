set -x
timeout 600 python -m pytest tests/test_conv_cl_gpu.py tests/test_kernels_gpu.py tests/test_train_gpu.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo pytest_conv=$?; tail -8 gpurun_out/pytest_conv.log | cut -c1-300
timeout 300 python tools/gpu_bench_layers.py 160 > gpurun_out/layers_r02d.log 2>&1; head -20 gpurun_out/layers_r02d.log
timeout 200 python tools/gpu_trace_conv.py > gpurun_out/trace_r02d.log 2>&1; head -140 gpurun_out/trace_r02d.log | cut -c1-150
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_all.log 2>&1; echo pytest_all=$?; tail -8 gpurun_out/pytest_all.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err; echo bench=$?; cut -c1-300 gpurun_out/bench_r02d.json
