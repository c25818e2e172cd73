"""Summarise `ncu --page raw --csv` exports: per launch duration, DRAM bytes, pipe utilisation, occupancy.
usage: python tools/roofline_from_ncu.py matvec_raw.csv [more_raw.csv ...]   (prints a markdown table; the first file's DRAM
bytes, summed over its launches, are what profiles/roofline_traffic.json records for one matvec)"""
import csv
import json
import sys

KEYS = [('gpu__time_duration.sum', 'us', 1e-3), ('dram__bytes_read.sum', 'MB read', 1e-6), ('dram__bytes_write.sum', 'MB written', 1e-6),
        ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'DMMA pipe %', 1.0),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'FP64 (DFMA) pipe %', 1.0),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %', 1.0),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %', 1.0),
        ('launch__registers_per_thread', 'regs', 1.0), ('launch__grid_size', 'grid', 1.0), ('launch__cluster_dim_x', 'cluster', 1.0)]


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return None


total = None
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path, errors='replace')))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, units = rows[hi], rows[hi + 1]
    print('### %s\n' % path)
    cols = [(k, lab, sc) for k, lab, sc in KEYS if k in hdr]
    print('| kernel | ' + ' | '.join(lab for _, lab, _ in cols) + ' |')
    print('|---|' + '---|' * len(cols))
    tot = 0.0
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')[:60]
        vals = []
        for k, lab, sc in cols:
            v = num(r[hdr.index(k)])
            u = units[hdr.index(k)]
            if v is not None and k.startswith('gpu__time'):
                v = v * {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'second': 1e6}.get(u, 1.0)
            elif v is not None and k.startswith('dram__bytes'):
                v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1e-6)
                tot += v
            vals.append('%.1f' % v if v is not None else '-')
        print('| `%s` | ' % name + ' | '.join(vals) + ' |')
    print()
    if total is None:
        total = tot
if total is not None:
    print('DRAM bytes of the first file, all launches: %.1f MB' % total)
    print('JSON: ' + json.dumps({'dram_bytes_first_file': total * 1e6}))
