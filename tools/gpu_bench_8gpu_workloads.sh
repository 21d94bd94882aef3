for wl in train inference train_c6; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${wl}_8gpu_r02.json 2> gpurun_out/bench8_$wl.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${wl}_8gpu_r02.json') if l.startswith('{')][-1]); print('N=8 $wl', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'), d['clocks']['reasons'])" || tail -20 gpurun_out/bench8_$wl.err
done
