mkdir -p gpurun_out/D
for div in 1 2; do for chi in 256 512; do
echo "div=$div" ; TNALG_OVERLAP_GRID_DIV=$div python tools/profile_lanczos.py --chi $chi --solves 10 2>&1 | tail -1
done; done
echo no-overlap; TNALG_NO_OVERLAP=1 python tools/profile_lanczos.py --chi 512 --solves 10 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:chain_gemm_tma -s 40 -c 2 -o gpurun_out/D/gemm512 -f python tools/profile_lanczos.py --chi 512 --solves 2 > /dev/null 2>&1
ncu -i gpurun_out/D/gemm512.ncu-rep --page raw --csv > gpurun_out/D/gemm512_raw.csv 2>/dev/null
ncu -i gpurun_out/D/gemm512.ncu-rep --page source --csv > gpurun_out/D/gemm512_src.csv 2>/dev/null
ls -la gpurun_out/D
