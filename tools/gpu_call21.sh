timeout 600 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02o.json 2> gpurun_out/bench_r02o.err; echo bench=$?; cut -c1-200 gpurun_out/bench_r02o.json; tail -2 gpurun_out/bench_r02o.err
