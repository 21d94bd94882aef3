for v in 0 1; do echo "== spin $v"; PGV_CL_SPIN=$v timeout 300 python tools/gpu_bench_layers.py 160 2>&1 | grep -E "^enc[2-7]|^enc8|totals"; done
