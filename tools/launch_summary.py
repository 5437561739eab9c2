"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
   python tools/launch_summary.py launches.csv > summary.txt"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
kn, mn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[mn] != 'gpu__time_duration.sum':
        continue
    v = float(r[mv].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[mu], 1.0)
    t = tot.setdefault(r[kn], [0, 0.0])
    t[0] += 1
    t[1] += v
allt = sum(t[1] for t in tot.values())
print('# per-kernel totals of %s (ncu gpu__time_duration.sum, cold-cache serialised: compare SHARES)' % sys.argv[1])
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-72s n=%4d total=%10.1f us avg=%9.2f us share=%5.1f%%' % (k[:72], n, t, t / n, 100 * t / allt))
