set -x
timeout 600 python -m pytest tests/test_conv_cl_gpu.py tests/test_metrics_gpu.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo pytest_conv=$?; tail -15 gpurun_out/pytest_conv.log
timeout 300 python tools/gpu_bench_layers.py 160 > gpurun_out/layers_r02b.log 2>&1; tail -28 gpurun_out/layers_r02b.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_all.log 2>&1; echo pytest_all=$?; tail -15 gpurun_out/pytest_all.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo bench=$?; cut -c1-300 gpurun_out/bench_r02b.json
