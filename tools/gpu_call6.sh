set -x
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_metrics_gpu.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo pytest_train=$?; tail -5 gpurun_out/pytest_conv.log | cut -c1-300
timeout 400 python tools/gpu_step_breakdown.py 160 > gpurun_out/breakdown_r02e.log 2>&1; cat gpurun_out/breakdown_r02e.log | cut -c1-150
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_train_r02e.csv python tools/run_train_once.py 160 2 > gpurun_out/ncu_l.log 2>&1; echo ncu=$?
python tools/summarize_launches.py gpurun_out/launches_train_r02e.csv gpurun_out/traffic_r02e.json > gpurun_out/launches_train_r02e.md; head -60 gpurun_out/launches_train_r02e.md
