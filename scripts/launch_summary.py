"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by
kernel name.  usage: launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3,
         'msecond': 1.0, 's': 1e3, 'second': 1e3}
for r in rows[1:]:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = r[ix['Kernel Name']].split('(')[0][-64:]
    v = float(r[ix['Metric Value']].replace(',', ''))
    agg[name][0] += 1
    agg[name][1] += v * scale[r[ix['Metric Unit']]]
tot = sum(v[1] for v in agg.values())
print("total kernel time %.2f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-66s n=%5d total=%9.2f ms avg=%8.3f ms  %5.1f%%'
          % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
