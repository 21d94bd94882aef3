timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_8gpu_r02c.json 2> gpurun_out/bench8.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_8gpu_r02c.json') if l.startswith('{')][-1]); print('N=8', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))" || tail -30 gpurun_out/bench8.err
