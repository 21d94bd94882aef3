set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytestD.log 2>&1; echo pytest=$?
grep -E "passed|failed|^FAILED" gpurun_out/pytestD.log | tail -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smokeD.log 2>&1; echo smoke=$?; tail -2 gpurun_out/smokeD.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_train_r01.json 2> gpurun_out/bench_train.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_train_r01.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d.get('ms_by_entry_point_eager'))"
timeout 300 python tools/gpu_determinism.py 16 > gpurun_out/determinism.log 2>&1; echo det=$?; tail -32 gpurun_out/determinism.log
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_train_r01.csv python tools/run_train_once.py 160 2 > gpurun_out/ncu_l.log 2>&1; echo ncu=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_cl_kernel -c 4 -f -o gpurun_out/conv_cl_r01 python tools/ncu_conv_probe.py > gpurun_out/ncu_f.log 2>&1; echo ncufull=$?
