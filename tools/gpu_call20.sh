timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02.log 2>&1; echo pytest=$?; tail -5 gpurun_out/pytest_gpu_r02.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err; echo bench=$?; cut -c1-200 gpurun_out/bench_r02n.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pipeline > gpurun_out/bench_r02n_nopipe.json 2> gpurun_out/bench_r02n_nopipe.err; echo bench_np=$?; cut -c1-200 gpurun_out/bench_r02n_nopipe.json
