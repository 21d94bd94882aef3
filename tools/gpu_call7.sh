set -x
nvidia-smi -L
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_all.log 2>&1; echo pytest_all=$?; tail -12 gpurun_out/pytest_all.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_r02a.json 2> gpurun_out/bench_2gpu_r02a.err; echo bench2=$?; cut -c1-260 gpurun_out/bench_2gpu_r02a.json; tail -5 gpurun_out/bench_2gpu_r02a.err | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; echo bench=$?; cut -c1-260 gpurun_out/bench_r02f.json
