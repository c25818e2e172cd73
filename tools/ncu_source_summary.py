"""Summarise an `ncu --page source --csv` export: stall-reason totals and opcode mix.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv; python tools/ncu_source_summary.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot, opc, opc_inst, n = collections.Counter(), collections.Counter(), collections.Counter(), 0
top = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[idx['# Samples']].isdigit():
        continue
    s = int(r[idx['# Samples']])
    n += s
    for c in stall_cols:
        tot[c] += int(r[idx[c]] or 0)
    toks = r[idx['Source']].split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '')
    opc[op] += s
    opc_inst[op] += int(r[idx['Instructions Executed']] or 0)
    top.append((s, r[idx['Source']].strip()))
print('total samples', n)
for k, v in tot.most_common(10):
    print('%-26s %8d %5.1f%%' % (k, v, 100.0 * v / max(n, 1)))
ti = sum(opc_inst.values())
print('--- opcode mix (stall samples | instructions executed)')
for k, v in sorted(opc_inst.items(), key=lambda kv: -kv[1])[:22]:
    print('%-22s samples %7d %5.1f%%   inst %12d %5.1f%%' % (k, opc[k], 100.0 * opc[k] / max(n, 1), v, 100.0 * v / ti))
print('--- hottest instructions')
for s, src in sorted(top, reverse=True)[:14]:
    print('%7d  %s' % (s, src[:110]))
