#!/bin/bash
# One-GPU evidence run (through gpurun): pytest -m gpu, the default bench line, QR timings vs cuSOLVER, two-site sweep profile.
# Outputs: gpurun_out/suite/.
mkdir -p gpurun_out/suite
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/suite/pytest_gpu.log 2>&1
tail -5 gpurun_out/suite/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/suite/bench_1gpu.json 2> gpurun_out/suite/bench_1gpu.err
tail -c 600 gpurun_out/suite/bench_1gpu.err
timeout 300 python tools/profile_qr.py > gpurun_out/suite/qr.txt 2>&1
cat gpurun_out/suite/qr.txt | tail -15
timeout 300 python tools/profile_two_site.py > gpurun_out/suite/two_site.txt 2>&1
tail -15 gpurun_out/suite/two_site.txt
