timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_train_2gpu_r01.json 2> gpurun_out/bench2.err; echo rc=$?
tail -c 1500 gpurun_out/bench_train_2gpu_r01.json | cut -c1-600; tail -5 gpurun_out/bench2.err
