mkdir -p gpurun_out/A
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/A/pytest_gpu.log 2>&1
tail -5 gpurun_out/A/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/A/bench_1gpu.json 2> gpurun_out/A/bench_1gpu.err
tail -c 600 gpurun_out/A/bench_1gpu.err
timeout 300 python tools/profile_qr.py > gpurun_out/A/qr.txt 2>&1
cat gpurun_out/A/qr.txt | tail -15
timeout 300 python tools/profile_two_site.py > gpurun_out/A/two_site.txt 2>&1
tail -15 gpurun_out/A/two_site.txt
