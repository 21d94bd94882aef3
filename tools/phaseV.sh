timeout 200 python tools/gpu_trace_conv.py > gpurun_out/trace_cl.log 2>&1; cat gpurun_out/trace_cl.log
