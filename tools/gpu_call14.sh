set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_2gpu_r02b.json 2> gpurun_out/bench2.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_2gpu_r02b.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))"; tail -3 gpurun_out/bench2.err
timeout 300 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err; echo bench=$?; cut -c1-200 gpurun_out/bench_r02l.json
