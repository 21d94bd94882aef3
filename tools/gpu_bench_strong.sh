# Strong scaling of the training step (global batch 160 split over N ranks): tools/gpurun_retry.sh --gpus N -- 'bash tools/gpu_bench_strong.sh N'
N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --strong-scaling --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_strong_${N}gpu_r02.json 2> gpurun_out/bench_strong$N.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_strong_${N}gpu_r02.json') if l.startswith('{')][-1]); print('N=$N strong', d['value'], d['ms_per_step'], d['config']['global_batch'], d['e2e']['value'])" || tail -20 gpurun_out/bench_strong$N.err
