set -x
timeout 300 python -m pytest tests/test_flow_fused_gpu.py -x -q -m gpu > gpurun_out/pytest_flow.log 2>&1; echo pytest_flow=$?; tail -8 gpurun_out/pytest_flow.log | cut -c1-400
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py tests/test_metrics_gpu.py tests/test_parity_default_gpu.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo pytest_sel=$?; tail -5 gpurun_out/pytest_conv.log | cut -c1-400
timeout 300 python tools/gpu_timeline.py 160 > gpurun_out/timeline_r02c.log 2>&1; cat gpurun_out/timeline_r02c.log | cut -c1-120
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err; echo bench=$?; cut -c1-200 gpurun_out/bench_r02k.json; tail -3 gpurun_out/bench_r02k.err
