"""Pivot an `ncu --csv --metrics ...` log (long format) into one row per launch / one row per kernel.
usage: python tools/ncu_pivot.py <log.csv> [--step-kernel conv1a_kernel --step 1] [--out summary.csv]"""
import argparse
import csv
import re
import sys
from collections import OrderedDict, defaultdict

ap = argparse.ArgumentParser()
ap.add_argument('log')
ap.add_argument('--step-kernel', default='conv1a_kernel', help='kernel whose launches mark the start of a step')
ap.add_argument('--step', type=int, default=1, help='which step (0-based occurrence of --step-kernel) to summarise')
ap.add_argument('--out')
a = ap.parse_args()

launches = OrderedDict()
with open(a.log, newline='') as f:
    rows = csv.reader(l for l in f if l.startswith('"'))
    hdr = next(rows)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        if len(r) != len(hdr) or r[0] == 'ID':
            continue
        d = launches.setdefault(int(r[ix['ID']]), {'name': r[ix['Kernel Name']], 'grid': r[ix['Grid Size']], 'block': r[ix['Block Size']]})
        try:
            v = float(r[ix['Metric Value']].replace(',', ''))
        except ValueError:
            continue
        unit = r[ix['Metric Unit']]
        name = r[ix['Metric Name']]
        if name == 'gpu__time_duration.sum':
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)  # -> us
        elif unit in ('Kbyte', 'Mbyte', 'Gbyte'):
            v *= {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
        d[name] = v


def short(n):
    n = re.sub(r'^void\s+', '', n)
    n = re.sub(r'\(.*$', '', n)
    n = re.sub(r'<unnamed>::', '', n)
    return n[:70]


ids = list(launches)
marks = [i for i in ids if a.step_kernel in launches[i]['name']]
if len(marks) > a.step:
    lo = marks[a.step]
    hi = marks[a.step + 1] if len(marks) > a.step + 1 else ids[-1] + 1
else:
    lo, hi = ids[0], ids[-1] + 1
sel = [launches[i] for i in ids if lo <= i < hi]
agg = defaultdict(lambda: defaultdict(float))
for d in sel:
    k = short(d['name'])
    g = agg[k]
    g['n'] += 1
    t = d.get('gpu__time_duration.sum', 0.0)
    g['us'] += t
    g['dram'] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    g['l2'] += d.get('lts__t_bytes.sum', 0.0)
    for m, key in (('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
                   ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pct'),
                   ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
                   ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_pct')):
        g[key] += d.get(m, 0.0) * t  # time-weighted
    g['regs'] = max(g['regs'], d.get('launch__registers_per_thread', 0.0))
total = sum(g['us'] for g in agg.values())
out = [['kernel', 'launches', 'us_total', 'share_pct', 'dram_MB', 'dram_GBps', 'dram_pct_of_peak', 'tensor_pipe_pct', 'sm_throughput_pct',
        'warps_active_pct', 'l2_MB', 'regs']]
for k, g in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    t = g['us'] or 1e-9
    out.append([k, int(g['n']), f"{g['us']:.1f}", f"{100 * g['us'] / total:.1f}", f"{g['dram'] / 1e6:.1f}", f"{g['dram'] / t / 1e3:.0f}",
                f"{g['dram_pct'] / t:.1f}", f"{g['tensor_pct'] / t:.1f}", f"{g['sm_pct'] / t:.1f}", f"{g['warps_pct'] / t:.1f}",
                f"{g['l2'] / 1e6:.1f}", int(g['regs'])])
w = csv.writer(open(a.out, 'w', newline='') if a.out else sys.stdout)
w.writerows(out)
print(f'# launches {lo}..{hi - 1} ({len(sel)} launches, {total / 1e3:.2f} ms of kernel time)', file=sys.stderr)
