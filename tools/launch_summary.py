"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
kn, mv, un = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(',', ''))
    except ValueError:
        continue
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[un].strip(), 1e-6)
    name = re.sub(r'\(.*', '', r[kn])
    name = re.sub(r'^void ', '', name)[:100]
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print('| kernel | launches | total ms | share |\n|---|---|---|---|')
for k, v in tot.most_common(25):
    print('| `%s` | %d | %.2f | %.1f %% |' % (k, cnt[k], v, 100 * v / total))
print('| **all** | %d | %.2f | 100 %% |' % (sum(cnt.values()), total))
