"""Top stalled SASS lines + stall-reason totals of an ncu source-page CSV
(ncu -i rep --page source --csv --print-source sass)."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot, 'instructions', len(data))
c = Counter(); n = 0
for r in data:
    src = r[ix['Source']]
    if 'BRA' in src and int(r[ix['stall_long_sb']] or 0) > 0.9 * int(r[ix['# Samples']] or 1):
        continue   # mbarrier spin loops
    for h in stalls:
        c[h] += int(r[ix[h]] or 0)
    n += int(r[ix['# Samples']] or 0)
print('non-spin samples', n)
for h, v in c.most_common(8):
    print('  %-24s %8d %5.1f%%' % (h, v, 100 * v / n))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:ntop]:
    s = int(r[ix['# Samples']])
    st = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print('%5d %-64s %7d %4.1f%% x%s %s' % (data.index(r), r[ix['Source']][:64], s, 100 * s / tot,
                                         r[ix['Instructions Executed']], st))
