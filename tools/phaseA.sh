set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest11.log 2>&1; echo pytest=$?
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke11.log 2>&1; echo smoke=$?
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_train_r01.json 2> gpurun_out/bench_train.err; echo bench=$?
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_train_r01.csv python tools/run_train_once.py 160 2 > gpurun_out/ncu_l.log 2>&1; echo ncu=$?
tail -3 gpurun_out/pytest11.log; tail -3 gpurun_out/smoke11.log; cat gpurun_out/bench_train_r01.json | cut -c1-600
