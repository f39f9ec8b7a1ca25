"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[row['Metric Unit']]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print('%-64s %8s %12s %7s' % ('kernel', 'launches', 'total ms', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-64s %8d %12.3f %6.1f%%' % (k[:64], v[0], v[1], 100 * v[1] / tot))
    print('%-64s %8d %12.3f' % ('TOTAL', sum(v[0] for v in agg.values()), tot))


if __name__ == '__main__':
    main(sys.argv[1])
