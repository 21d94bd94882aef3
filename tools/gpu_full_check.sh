# Round-end verification on one B200 (run through gpurun): GPU tests, smoke, ncu launch list + traffic, the four bench lines,
# one ncu --set full capture of conv_cl_kernel, per-layer / small-kernel / pipeline-trace logs.  Outputs land in gpurun_out/.
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytestZ.log 2>&1; echo pytest=$?; tail -2 gpurun_out/pytestZ.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smokeZ.log 2>&1; echo smoke=$?; tail -1 gpurun_out/smokeZ.log
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_train_r01.csv python tools/run_train_once.py 160 2 > gpurun_out/ncu_l.log 2>&1; echo ncu=$?
python tools/summarize_launches.py gpurun_out/launches_train_r01.csv profiles/traffic_r01.json > gpurun_out/launches_train_r01.md
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_train_r01.json 2> gpurun_out/bench_train.err; echo bench=$?
timeout 300 python bench.py --workload frontend --steps 20 --warmup 5 > gpurun_out/bench_frontend_r01.json 2>> gpurun_out/bench_train.err; echo benchf=$?
timeout 300 python bench.py --workload inference --steps 20 --warmup 5 > gpurun_out/bench_inference_r01.json 2>> gpurun_out/bench_train.err; echo benchi=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_r01.json 2>> gpurun_out/bench_train.err; echo benchr=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_cl_kernel -c 4 -f -o gpurun_out/conv_cl_r01 python tools/ncu_conv_probe.py > gpurun_out/ncu_f.log 2>&1; echo ncufull=$?
timeout 300 python tools/gpu_bench_layers.py 160 > gpurun_out/layers_M.log 2>&1
timeout 300 python tools/gpu_bench_small.py 160 > gpurun_out/small_M.log 2>&1
timeout 200 python tools/gpu_trace_conv.py > gpurun_out/trace_cl.log 2>&1
cp profiles/traffic_r01.json gpurun_out/traffic_r01.json
for f in gpurun_out/bench_train_r01.json gpurun_out/bench_frontend_r01.json gpurun_out/bench_inference_r01.json gpurun_out/bench_reference_r01.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), (d.get('roofline') or {}).get('kernel'), (d.get('roofline') or {}).get('frac'))"; done
