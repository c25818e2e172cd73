mkdir -p gpurun_out/C
for chi in 256 512; do
python tools/profile_lanczos.py --chi $chi > gpurun_out/C/lanczos_wall_$chi.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/C/lz_launches_$chi.csv python tools/profile_lanczos.py --chi $chi --solves 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/C/lz_launches_$chi.csv > gpurun_out/C/lz_launches_$chi.md
cat gpurun_out/C/lanczos_wall_$chi.txt gpurun_out/C/lz_launches_$chi.md
done
