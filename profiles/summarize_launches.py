#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share of the last step."""
import collections
import csv
import re
import sys


def main(path, steps_total):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', ''))) for x in csv.DictReader(lines)]
    per = len(rows) // steps_total
    last = rows[-per:]
    tot = sum(v for _, v in last)
    agg = collections.OrderedDict()
    for k, v in last:
        k = re.sub(r'\(.*', '', k)
        agg.setdefault(k, [0, 0])
        agg[k][0] += v
        agg[k][1] += 1
    print('# %s: %d launches total, %d per step; last step' % (path, len(rows), per))
    for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print('%-58s n=%2d %10.1f us %5.1f%%' % (k[:58], c, v / 1e3, 100 * v / tot))
    print('total %.1f us' % (tot / 1e3))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]))
