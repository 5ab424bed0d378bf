#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import collections
import csv
import sys


def main(path, skip=0):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        n += 1
        if n <= skip:
            continue
        k = row['Kernel Name'].split('(')[0][-70:]
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print('%d launches, %.1f us total' % (n - skip, tot))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s n=%4d total=%10.1f us  avg=%9.1f us  share=%5.1f%%' % (k, c, t, t / c, 100 * t / tot))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
