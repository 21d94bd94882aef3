set -x
timeout 600 python -m pytest tests/test_conv_cl_gpu.py tests/test_train_gpu.py tests/test_multigpu_gpu.py tests/test_flow_fused_gpu.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo pytest_sel=$?; tail -5 gpurun_out/pytest_conv.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; echo bench=$?; cut -c1-200 gpurun_out/bench_r02g.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_r02b.json 2> gpurun_out/bench_2gpu_r02b.err; echo bench2=$?; cut -c1-200 gpurun_out/bench_2gpu_r02b.json
NCCL_MAX_CTAS=8 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_r02b_cta8.json 2> gpurun_out/bench_2gpu_r02b_cta8.err; echo bench2c=$?; cut -c1-200 gpurun_out/bench_2gpu_r02b_cta8.json
timeout 300 python tools/gpu_bench_layers.py 160 > gpurun_out/layers_r02g.log 2>&1; head -20 gpurun_out/layers_r02g.log
