"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST step."""
import collections
import csv
import re
import sys


def ms(row):
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    return v / 1e6 if u.startswith('n') else v / 1e3 if u.startswith('u') else v


def main(path, marker='patchify'):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    idx = [i for i, r in enumerate(rows) if marker in r['Kernel Name']]
    step = rows[idx[-1]:]
    tot = sum(ms(r) for r in step)
    agg = collections.OrderedDict()
    for r in step:
        n = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('clipn::', '')
        n = re.sub(r'at::.*?::', 'at::', n)[:60]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms(r)
    print(f"last step: {len(step)} launches, {tot:.2f} ms (serialised, cold cache)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f"{t:8.2f} ms {100 * t / tot:5.1f}%  n={c:4d}  avg {1e3 * t / c:7.1f} us  {k}")


if __name__ == '__main__':
    main(*sys.argv[1:])
