"""CPU: the bench.py contract that can be checked without a GPU - the reference arm runs here (it is the CPU oracle port), prints ONE
JSON line with the agreed keys, and describes its workload with exactly the `config` object our arm prints for the same flags."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line_and_shared_config():
    import bench
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--batch-per-gpu', '2', '--steps', '1',
                        '--warmup', '1'], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['metric'] == bench.METRIC['train'] and d['unit'] == 'samples/s' and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config'] == bench.line_config('train', 2, 1, True)          # what our arm prints for the same flags
    assert 'model' not in d['config'] and 'workload' in d['config']


def test_workloads_cover_the_baseline_configs():
    import bench
    assert set(bench.METRIC) >= {'train', 'train_c6', 'frontend', 'inference'}
    assert bench.default_batch('train') == 160 and bench.default_batch('train_c6') == 160 and bench.default_batch('inference') == 1024
    for w in bench.METRIC:
        assert str(bench.default_batch(w)) in bench.workload_name(w, bench.default_batch(w))
