set -x
timeout 300 python -m pytest tests/test_flow_fused_gpu.py -x -q -m gpu > gpurun_out/pytest_flow.log 2>&1; echo pytest_flow=$?; tail -8 gpurun_out/pytest_flow.log | cut -c1-400
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py tests/test_metrics_gpu.py tests/test_parity_default_gpu.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo pytest_sel=$?; tail -5 gpurun_out/pytest_conv.log | cut -c1-400
timeout 300 python tools/gpu_timeline.py 160 > gpurun_out/timeline_r02b.log 2>&1; cat gpurun_out/timeline_r02b.log | cut -c1-120
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err; echo bench=$?; cut -c1-200 gpurun_out/bench_r02j.json; tail -3 gpurun_out/bench_r02j.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pipeline > gpurun_out/bench_r02j_nopipe.json 2> gpurun_out/bench_r02j_nopipe.err; echo bench_np=$?; cut -c1-200 gpurun_out/bench_r02j_nopipe.json
