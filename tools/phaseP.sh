for g in 592 1184 296 592 1184; do PGV_BN_GRID=$g PGV_OVERLAP_WGRAD=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bn_grid=$g', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"; done
