"""Summarises an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel launch count, summed duration, share of the step and DRAM traffic for the LAST training step in the file
(a step starts at the front end's split_audio_kernel).  Usage: summarize_launches.py launches.csv > summary.md"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r'^void\s+', '', name)
    m = re.match(r'(?:pgv::)?(\w+)(<[^(]*>)?', name)
    base = m.group(1) if m else name[:40]
    tmpl = m.group(2) or '' if m else ''
    if name.startswith('at::') or 'at::native' in name:
        return 'ATen: ' + re.sub(r'\(.*', '', name)[:70]
    tmpl = re.sub(r'pgv::', '', tmpl)
    return base + (tmpl if len(tmpl) < 30 else '')


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    launches = collections.OrderedDict()
    for row in rd:
        rec = launches.setdefault(int(row['ID']), {'name': row['Kernel Name']})
        rec[row['Metric Name']] = float(row['Metric Value'].replace(',', ''))
        rec['unit_' + row['Metric Name']] = row['Metric Unit']
    ids = sorted(launches)
    starts = [i for i in ids if 'split_audio_kernel' in launches[i]['name']]
    first = starts[-1] if starts else ids[0]
    step = [launches[i] for i in ids if i >= first]
    agg = collections.OrderedDict()
    for rec in step:
        a = agg.setdefault(short(rec['name']), [0, 0.0, 0.0, 0.0])
        t = rec.get('gpu__time_duration.sum', 0.0)
        if rec.get('unit_gpu__time_duration.sum', 'ns') in ('ns', 'nsecond'):
            t /= 1e3        # -> us
        a[0] += 1; a[1] += t
        a[2] += rec.get('dram__bytes_read.sum', 0.0); a[3] += rec.get('dram__bytes_write.sum', 0.0)
    tot = sum(a[1] for a in agg.values())
    print('launches in the last step: %d, summed kernel time %.2f ms (serialised, cold caches: shares matter, not absolutes)\n' % (len(step), tot / 1e3))
    print('| kernel | launches | time (us) | share | DRAM read (MB) | DRAM write (MB) |')
    print('|---|---|---|---|---|---|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| %s | %d | %.1f | %.1f%% | %.1f | %.1f |' % (k, a[0], a[1], 100 * a[1] / tot, a[2] / 1e6, a[3] / 1e6))
    if len(sys.argv) > 2:        # per-kernel-family DRAM traffic of one step, for bench.py's roofline.traffic
        import json
        fam = {}
        for k, a in agg.items():
            f = fam.setdefault(re.sub(r'<.*', '', k), dict(launches=0, us=0.0, dram_bytes=0.0))
            f['launches'] += a[0]; f['us'] += a[1]; f['dram_bytes'] += a[2] + a[3]
        json.dump(fam, open(sys.argv[2], 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main(sys.argv[1])
