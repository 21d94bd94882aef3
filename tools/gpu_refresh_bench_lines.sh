R=r02
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$R.log 2>&1; echo pytest=$?; tail -2 gpurun_out/pytest_$R.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo smoke=$?; tail -1 gpurun_out/smoke_$R.log
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_train_$R.csv python tools/run_train_once.py 160 2 > gpurun_out/ncu_l.log 2>&1; echo ncu=$?
python tools/summarize_launches.py gpurun_out/launches_train_$R.csv profiles/traffic_$R.json > gpurun_out/launches_train_$R.md
cp profiles/traffic_$R.json gpurun_out/traffic_$R.json
timeout 400 python bench.py > gpurun_out/bench_train_$R.json 2> gpurun_out/bench_train.err; echo bench=$?
timeout 300 python bench.py --workload frontend --steps 20 --warmup 5 > gpurun_out/bench_frontend_$R.json 2>> gpurun_out/bench_train.err; echo benchf=$?
timeout 300 python bench.py --workload inference --steps 20 --warmup 5 > gpurun_out/bench_inference_$R.json 2>> gpurun_out/bench_train.err; echo benchi=$?
timeout 300 python bench.py --workload train_c6 --steps 10 --warmup 3 > gpurun_out/bench_train_c6_$R.json 2>> gpurun_out/bench_train.err; echo benchc6=$?
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$R.json 2>> gpurun_out/bench_train.err; echo benchr=$?
timeout 200 python tools/gpu_timeline.py 160 > gpurun_out/timeline_$R.log 2>&1
timeout 200 python tools/gpu_step_breakdown.py 160 > gpurun_out/breakdown_$R.log 2>&1
for f in train frontend inference train_c6 reference; do python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${f}_$R.json')); print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), (d.get('roofline') or {}).get('kernel'), (d.get('roofline') or {}).get('frac'), (d.get('clocks') or {}).get('reasons'))"; done
