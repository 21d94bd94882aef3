for env in "NCCL_NVLS_ENABLE=0" "NCCL_ALGO=Ring" "NCCL_PROTO=Simple"; do
env $env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nccl_try.json 2> gpurun_out/bench_nccl_try.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_nccl_try.json') if l.startswith('{')][-1]); print('$env', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/bench_nccl_try.err
done
