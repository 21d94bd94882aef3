set -x
timeout 600 python -m pytest tests/test_flow_fused_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py tests/test_metrics_gpu.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo pytest_sel=$?; tail -5 gpurun_out/pytest_conv.log | cut -c1-400
timeout 300 python tools/gpu_timeline.py 160 > gpurun_out/timeline_r02.log 2>&1; cat gpurun_out/timeline_r02.log | cut -c1-120
PGV_PDL=0 timeout 300 python tools/gpu_timeline.py 160 > gpurun_out/timeline_r02_nopdl.log 2>&1; head -3 gpurun_out/timeline_r02_nopdl.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err; echo bench=$?; cut -c1-200 gpurun_out/bench_r02i.json; tail -3 gpurun_out/bench_r02i.err
PGV_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02i_nopdl.json 2> gpurun_out/bench_r02i_nopdl.err; echo bench_nopdl=$?; cut -c1-200 gpurun_out/bench_r02i_nopdl.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pipeline > gpurun_out/bench_r02i_nopipe.json 2> gpurun_out/bench_r02i_nopipe.err; echo bench_np=$?; cut -c1-200 gpurun_out/bench_r02i_nopipe.json
