set -x
for mode in 0 1; do
PGV_DEBUG_SKIP_ALLREDUCE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_2gpu_skip$mode.json 2> gpurun_out/bench2.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_2gpu_skip$mode.json') if l.startswith('{')][-1]); print('skip_allreduce=$mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))"
done
timeout 300 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -3
