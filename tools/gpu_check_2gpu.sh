# Two GPUs (gpurun --gpus 2): external-event semantics, overlapped vs plain all-reduce, 2-GPU train bench.
timeout 120 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k external 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_overlap_allreduce.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_2gpu_r02.json 2> gpurun_out/bench2.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_2gpu_r02.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -3 gpurun_out/bench2.err
