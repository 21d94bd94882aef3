"""Markdown table of the headline metrics of every launch in an .ncu-rep (read through `ncu -i ... --page raw --csv`).
Usage: ncu_summary.py report.ncu-rep "title" [label1 label2 ...] > summary.md"""
import csv
import io
import subprocess
import sys

ROWS = [('duration (us)', 'gpu__time_duration.sum', 1e-3),
        ('grid / block', None, None),
        ('registers/thread', 'launch__registers_per_thread', 1),
        ('dynamic + static smem per CTA (KB)', None, None),
        ('DRAM read (MB)', 'dram__bytes_read.sum', None),
        ('DRAM write (MB)', 'dram__bytes_write.sum', None),
        ('DRAM throughput (% of peak)', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1),
        ('L2 throughput (%)', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 1),
        ('L1/TEX throughput (%)', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 1),
        ('SM throughput (%)', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 1),
        ('issue slots busy (%)', 'sm__inst_issued.avg.pct_of_peak_sustained_active', 1),
        ('tensor pipe active (% of active cycles)', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 1),
        ('utcmma tf32 ops (% of peak, elapsed)', 'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 1),
        ('TMEM pipe instructions (%)', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active', 1),
        ('fma pipe (%)', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 1),
        ('warps active (% of max)', 'sm__warps_active.avg.pct_of_peak_sustained_active', 1),
        ('executed warp instructions', 'smsp__inst_executed.sum', 1),
        ('stall: long scoreboard (warps/issue)', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 1),
        ('stall: short scoreboard', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 1),
        ('stall: barrier', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 1),
        ('stall: wait', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 1),
        ('stall: sleeping / membar', None, None)]


def main():
    rep, title, labels = sys.argv[1], sys.argv[2], sys.argv[3:]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=None):
        if name not in col:
            return 'n/a'
        v, u = r[col[name]].replace(',', ''), units[col[name]]
        try:
            f = float(v)
        except ValueError:
            return v
        if scale is None:                      # bytes -> MB
            f *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
        elif name == 'gpu__time_duration.sum':
            f *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(u, 1.0)
        return '%.4g' % f

    print('# %s\n' % title)
    print('Read with `ncu -i %s --page raw --csv` (tools/ncu_summary.py).  Times under ncu are cold-cache and serialised.\n' % rep.split('/')[-1])
    names = [(labels[i] if i < len(labels) else '') for i in range(len(data))]
    print('| metric | ' + ' | '.join('%s `%s`' % (n, r[col['Kernel Name']].split('(')[0].replace('void ', '').replace('pgv::', '')[:44]) for n, r in zip(names, data)) + ' |')
    print('|---|' + '---|' * len(data))
    for label, metric, scale in ROWS:
        cells = []
        for r in data:
            if label == 'grid / block':
                cells.append('%s / %s' % (val(r, 'launch__grid_size', 1), val(r, 'launch__block_size', 1)))
            elif label.startswith('dynamic'):
                d, s = r[col['launch__shared_mem_per_block_dynamic']], r[col['launch__shared_mem_per_block_static']]
                cells.append('%s + %s %s' % (d, s, units[col['launch__shared_mem_per_block_dynamic']]))
            elif label.startswith('stall: sleeping'):
                cells.append('%s / %s' % (val(r, 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 1),
                                          val(r, 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 1)))
            else:
                cells.append(val(r, metric, scale))
        print('| %s | ' % label + ' | '.join(cells) + ' |')


main()
