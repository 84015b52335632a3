"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, max, share."""
import collections
import csv
import re
import sys


def main(path, top=25, big=3.0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    order = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[row['Metric Unit']]
        short = re.sub(r'\(.*', '', row['Kernel Name'])
        short = re.sub(r'<unnamed>::|void ', '', short)[:60]
        a = agg[short]
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
        order.append((short, v))
    tot = sum(a[1] for a in agg.values())
    print('total kernel time %.3f ms over %d launches' % (tot, len(order)))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-60s n=%5d tot=%9.3f ms max=%8.3f share=%.3f' % (k, a[0], a[1], a[2], a[1] / tot))
    print('launches > %.1f ms:' % big, [(k[:18], round(v, 2)) for k, v in order if v > big])


if __name__ == '__main__':
    main(sys.argv[1])
