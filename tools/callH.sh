timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3
python tools/profile_qr.py 2>&1 | tail -10
