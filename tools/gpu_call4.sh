set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_all.log 2>&1; echo pytest_all=$?; tail -8 gpurun_out/pytest_all.log | cut -c1-300
timeout 300 python tools/gpu_bench_layers.py 160 > gpurun_out/layers_r02c.log 2>&1; head -20 gpurun_out/layers_r02c.log
timeout 300 python tools/gpu_bench_layers.py 160 cl noa > gpurun_out/layers_r02c_noa.log 2>&1; head -20 gpurun_out/layers_r02c_noa.log
timeout 200 python tools/gpu_trace_conv.py > gpurun_out/trace_r02c.log 2>&1; head -150 gpurun_out/trace_r02c.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo bench=$?; cut -c1-300 gpurun_out/bench_r02c.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_cl_kernel --launch-skip 3 --launch-count 3 -f -o gpurun_out/conv_thin_r02 python tools/ncu_conv_thin_probe.py > gpurun_out/ncu_thin.log 2>&1; echo ncu=$?
