"""Summarise ncu outputs into markdown (run here, no GPU):
    python tools/ncu_summary.py launches <launches.csv>         # per-kernel time shares
    python tools/ncu_summary.py full <report.ncu-rep>           # key metrics per captured launch
"""
import csv, io, re, subprocess, sys
from collections import OrderedDict

def short(n):
    n = re.sub(r'\(.*', '', n)
    n = n.replace('void ', '').replace('mvb::', '')
    return n.strip()

def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(io.StringIO(''.join(lines))))
    agg = OrderedDict()
    for r in rows:
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        k = short(r['Kernel Name'])
        v = float(r['Metric Value'].replace(',', ''))
        v = v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r['Metric Unit'], 1e-6)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f'total kernel time {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches\n')
    print('| kernel | launches | ms | share |\n|---|---|---|---|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {a[0]} | {a[1]:.2f} | {100 * a[1] / tot:.1f}% |')

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'lts__t_sector_hit_rate.pct']

def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"### launch {r[idx['ID']]}: `{short(r[idx['Kernel Name']])}` grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for k in KEYS:
            if k in idx:
                print(f'- {k}: {r[idx[k]]} {units[idx[k]]}')
        print()

if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
