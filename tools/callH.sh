timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -8
python tools/profile_qr.py 2>&1 | tail -10
echo "no fuse"; TNALG_QR_NO_FUSE=1 python tools/profile_qr.py 2>&1 | grep -E "^\| (512 x 256|1024 x 512|2048 x 1024|512 x 512) "
