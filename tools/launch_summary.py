"""Per-kernel totals of an ncu launch list (``--metrics gpu__time_duration.sum --csv``): python tools/launch_summary.py file.csv [skip]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1 + skip:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    n = r[ki].split('(')[0][:64]
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print('%-66s %5d %10.1f us %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
print('total %.1f us' % tot)
