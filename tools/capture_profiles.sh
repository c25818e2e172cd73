#!/bin/bash
# Round-2 evidence for profiles/ (run on the GPU box through gpurun; outputs land in gpurun_out/prof_r02/):
#   1. ncu launch list of the timed sweep of the default bench command (kernel shares)
#   2. ncu --set full of the two chain-GEMM launches of one matvec at the widest cfg4 site (DRAM traffic, DMMA pipe utilisation)
#   3. ncu --set full of the QR panel / apply kernels at 2048 x 1024
set -x
# every ncu step runs under its own `timeout`: on 2026-10-17 step 1 stalled inside ncu on the first chain_gemm_tma_kernel of the
# NVTX range and ate the rest of the round's GPU budget (the same command had run to completion earlier in the round)
OUT=gpurun_out/prof_r02
mkdir -p $OUT
# 1. launch list: 3000 consecutive launches (12000 took > 25 min under ncu) inside the timed sweep (nvtx range 'timed')
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 3000 --csv --log-file $OUT/launches_cfg4.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-other > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err
python tools/launch_summary.py $OUT/launches_cfg4.csv > $OUT/launches_cfg4.md
gzip -f $OUT/launches_cfg4.csv
# 2. matvec kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_gemm_tma -s 4 -c 2 -o $OUT/matvec -f \
    python tools/profile_matvec.py --chi 1024 --kl 4 --kr 4 --nx 15 --reps 2 > $OUT/matvec_run.txt 2>&1
ncu -i $OUT/matvec.ncu-rep --page raw --csv > $OUT/matvec_raw.csv
# 3. QR kernels
timeout 300 ncu --set full --clock-control none -k regex:"qr_panel|qr_apply" -s 150 -c 4 -o $OUT/qr -f python tools/qr_one.py > $OUT/qr_run.txt 2>&1
ncu -i $OUT/qr.ncu-rep --page raw --csv > $OUT/qr_raw.csv
python tools/roofline_from_ncu.py $OUT/matvec_raw.csv $OUT/qr_raw.csv > $OUT/ncu_summary.md
cat $OUT/ncu_summary.md
