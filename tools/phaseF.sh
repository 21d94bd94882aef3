set -x
timeout 300 python tools/gpu_bench_small.py 160 > gpurun_out/small_F.log 2>&1; cat gpurun_out/small_F.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:colslice_gemm_kernel -s 20 -c 3 -f -o gpurun_out/colslice_r01 python tools/gpu_bench_small.py 160 > gpurun_out/ncu_cs.log 2>&1; echo ncu=$?
