for i in 1 2 3; do python tools/profile_qr.py 2>&1 | grep -E "^\| (1024 x 512|2048 x 1024|2048 x 2048) "; done
