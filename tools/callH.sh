for r in 256 128 64; do echo "apply_rows=$r"; TNALG_QR_APPLY_CTA_ROWS=$r python tools/profile_qr.py 2>&1 | grep -E "^\| (512 x 256|1024 x 512|2048 x 1024|512 x 512|4096 x 2048) "; done
