#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum per launch): per-kernel totals and one epoch's sequence."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    return data


def short(n):
    n = n.replace('void ', '')
    return n[:n.index('(')][:46] if '(' in n else n[:46]


def main():
    path = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    data = load(path)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        a = agg[short(d['Kernel Name'])]
        a[0] += 1
        a[1] += float(d['Metric Value']) / 1e3
    tot = sum(a[1] for a in agg.values())
    names = [short(d['Kernel Name']) for d in data]
    idx = [i for i, n in enumerate(names) if n.startswith('k_basic<float, 0>') or n.startswith('k_perm_prep')]
    nsteps = max(len(idx), 1)
    print(f'{len(data)} launches, total {tot:.1f} us, {nsteps} steps, {tot / nsteps:.1f} us of kernel time per step')
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'{k:48s} n={a[0]:4d} us={a[1]:8.1f} share={a[1] / tot:.3f} avg={a[1] / a[0]:.1f}')
    if len(idx) > step:
        s = idx[step]
        e = idx[step + 1] if step + 1 < len(idx) else len(data)
        print(f'--- step {step} ---')
        for d in data[s:e]:
            print(short(d['Kernel Name']).ljust(48), '%7.1f' % (float(d['Metric Value']) / 1e3), d['Grid Size'], d['Stream'])


if __name__ == '__main__':
    main()
