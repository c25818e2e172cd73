#!/bin/bash
# one full ncu capture of the QR panel kernel (source-level stall reasons)
ncu --set full --clock-control none --import-source on -k regex:qr_panel -s 40 -c 1 -o gpurun_out/r2f/qr_panel -f python tools/qr_one.py > gpurun_out/r2f/ncu_panel.log 2>&1
ncu -i gpurun_out/r2f/qr_panel.ncu-rep --page source --csv > gpurun_out/r2f/qr_panel_src.csv 2>/dev/null
ncu -i gpurun_out/r2f/qr_panel.ncu-rep --page raw --csv > gpurun_out/r2f/qr_panel_raw.csv 2>/dev/null
