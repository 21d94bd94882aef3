timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02.log 2>&1; echo pytest=$?; tail -2 gpurun_out/pytest_r02.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; echo smoke=$?; tail -1 gpurun_out/smoke_r02.log
timeout 400 python bench.py > gpurun_out/bench_train_r02.json 2> gpurun_out/bench_train.err; echo bench=$?
timeout 300 python bench.py --workload inference --steps 20 --warmup 5 > gpurun_out/bench_inference_r02.json 2>> gpurun_out/bench_train.err; echo benchi=$?
python -c "
import json
for f in ('train','inference'):
    d=json.load(open('gpurun_out/bench_%s_r02.json'%f)); r=d['roofline']; print(f, d['value'], d['ms_per_step'], d['e2e']['value'], r['bound'], r['achieved'], r['peak'], r['frac'], r['frac_tensor'], r['frac_hbm'])"
