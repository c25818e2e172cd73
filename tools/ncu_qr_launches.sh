#!/bin/bash
# ncu launch lists (gpu__time_duration) of one Householder QR at 512x256 and 2048x1024 (tools/qr_one.py, NVTX ranges).
# Outputs: gpurun_out/qr_launches/.
mkdir -p gpurun_out/qr_launches
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "qr_512x256/" --csv --log-file gpurun_out/qr_launches/qr512.csv python tools/qr_one.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "qr_2048x1024/" --csv --log-file gpurun_out/qr_launches/qr2048.csv python tools/qr_one.py > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/qr_launches/qr512.csv; python tools/launch_summary.py gpurun_out/qr_launches/qr2048.csv
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/qr_launches/qr512.csv',errors='replace')))
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r); hdr=rows[hi]
kn,mv=hdr.index('Kernel Name'),hdr.index('Metric Value')
print([ (r[kn].split('(')[0].replace('void tn::','')[:28], r[mv]) for r in rows[hi+1:]])
P
