#!/bin/bash
# launch list + full captures of the chi = 256 Lanczos kernels (fused re-orthogonalisation, 64x64-tile GEMM stages)
mkdir -p gpurun_out/r2p
python tools/profile_lanczos.py --chi 256 > gpurun_out/r2p/lanczos_wall.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 120 --csv --log-file gpurun_out/r2p/lz_launches.csv python tools/profile_lanczos.py --chi 256 --solves 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lanczos_orth_fused -s 30 -c 1 -o gpurun_out/r2p/orth -f python tools/profile_lanczos.py --chi 256 --solves 2 > /dev/null 2>&1
ncu -i gpurun_out/r2p/orth.ncu-rep --page source --csv > gpurun_out/r2p/orth_src.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:chain_gemm_kernel -s 60 -c 2 -o gpurun_out/r2p/gemm -f python tools/profile_lanczos.py --chi 256 --solves 2 > /dev/null 2>&1
ncu -i gpurun_out/r2p/gemm.ncu-rep --page source --csv > gpurun_out/r2p/gemm_src.csv 2>/dev/null
ncu -i gpurun_out/r2p/gemm.ncu-rep --page raw --csv > gpurun_out/r2p/gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/r2p/orth.ncu-rep --page raw --csv > gpurun_out/r2p/orth_raw.csv 2>/dev/null
cat gpurun_out/r2p/lanczos_wall.txt
