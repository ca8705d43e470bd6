"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per kernel launches, total and average duration, share of the total."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[hdr]
    ki, vi = H.index('Kernel Name'), H.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 2:]:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0][:48]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{k:48s} n={v[0]:4d} total={v[1] / 1e3:10.1f}us '
              f'avg={v[1] / v[0] / 1e3:8.1f}us {100 * v[1] / tot:5.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
