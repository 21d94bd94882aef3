for v in 2 5 6; do echo "== variant $v"; PGV_CONV_VARIANT=$v timeout 300 python tools/gpu_bench_layers.py 160 2>&1 | grep -E "^enc[2-7]|^enc8" | awk '{print $1, $2, $3, $4, $5, $6, $7}'; done
