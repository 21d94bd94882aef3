N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_${N}gpu_r02.json 2> gpurun_out/bench$N.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_${N}gpu_r02.json') if l.startswith('{')][-1]); print('N=$N train', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'), d['clocks']['reasons'])" || tail -20 gpurun_out/bench$N.err
timeout 300 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -2
