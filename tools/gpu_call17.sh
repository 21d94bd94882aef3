for cta in 8 12 24; do
NCCL_MAX_CTAS=$cta timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_2gpu_cta$cta.json 2> gpurun_out/bench2.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_2gpu_cta$cta.json') if l.startswith('{')][-1]); print('NCCL_MAX_CTAS=$cta', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))"
done
