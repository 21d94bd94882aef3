timeout 300 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_2gpu_r02f.json 2> gpurun_out/bench2.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_2gpu_r02f.json') if l.startswith('{')][-1]); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))" || tail -30 gpurun_out/bench2.err
NCCL_MAX_CTAS=16 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_2gpu_r02f16.json 2> gpurun_out/bench2b.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_2gpu_r02f16.json') if l.startswith('{')][-1]); print('N=2 cta16', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/gpu_timeline_ddp.py > gpurun_out/timeline_ddp.log 2>&1; grep -v "^\*\|OMP_NUM" gpurun_out/timeline_ddp.log | tail -24
