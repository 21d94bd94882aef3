set -x
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo pytest_sel=$?; tail -5 gpurun_out/pytest_conv.log | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err; echo bench=$?; cut -c1-200 gpurun_out/bench_r02h.json; tail -3 gpurun_out/bench_r02h.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pipeline > gpurun_out/bench_r02h_nopipe.json 2> gpurun_out/bench_r02h_nopipe.err; echo bench_np=$?; cut -c1-200 gpurun_out/bench_r02h_nopipe.json
