set -x
timeout 900 python -m pytest tests/test_conv_cl_gpu.py tests/test_kernels_gpu.py -q -s -k "channels_last or cl or tensor_core_convs or thin or layout or repacking or block_forward" > gpurun_out/pytestB.log 2>&1; echo pytest=$?
grep -E "passed|failed|tc conv" gpurun_out/pytestB.log | tail -40
timeout 300 python tools/gpu_bench_layers.py 160 > gpurun_out/layers_cl.log 2>&1; echo layers=$?; cat gpurun_out/layers_cl.log
timeout 300 python tools/gpu_determinism.py 16 > gpurun_out/determinism.log 2>&1; echo det=$?; tail -40 gpurun_out/determinism.log
