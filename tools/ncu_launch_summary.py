#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,...` launch list: the last full iteration launch by launch and
per-kernel totals.   python tools/ncu_launch_summary.py gpurun_out/launches.csv [n_last]"""
import csv
import sys
from collections import OrderedDict, defaultdict


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    idx = {k: i for i, k in enumerate(hdr)}
    recs = OrderedDict()
    for r in rows[1:]:
        d = recs.setdefault(r[idx['ID']], {'name': r[idx['Kernel Name']], 'grid': r[idx['Grid Size']], 'block': r[idx['Block Size']]})
        v = float(r[idx['Metric Value']].replace(',', ''))
        unit = r[idx['Metric Unit']]
        name = r[idx['Metric Name']]
        if name == 'gpu__time_duration.sum':
            v = v / 1e3 if unit in ('ns', 'nsecond') else v          # -> us
        d[name] = v
    return list(recs.values())


def short(n):
    n = n.replace('void ', '').replace('<unnamed>::', '')
    return n.split('(')[0][:44]


def main():
    L = load(sys.argv[1])
    n_last = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    print('# last %d launches' % n_last)
    for x in L[-n_last:]:
        print('%-44s grid %-13s %8.1f us  tensor %5.1f %%  dram %6.1f MB' % (
            short(x['name']), x['grid'], x['gpu__time_duration.sum'],
            x.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0),
            (x.get('dram__bytes_read.sum', 0) + x.get('dram__bytes_write.sum', 0)) / 1e6))
    tot = defaultdict(lambda: [0, 0.0])
    for x in L:
        t = tot[short(x['name'])]
        t[0] += 1
        t[1] += x['gpu__time_duration.sum']
    s = sum(v[1] for v in tot.values())
    print('# per kernel over %d launches (cold-cache, serialised: compare SHARES)' % len(L))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print('%-44s n %5d  total %9.1f us  avg %8.1f us  share %5.1f %%' % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / s))


if __name__ == '__main__':
    main()
