"""Timing of tn_svd_jacobi vs cuSOLVER (torch.linalg.svd) and of the QR used by gauge moves."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnalg_b200 import ops  # noqa: E402

be = ops.backend()
g = torch.Generator(device=be.device).manual_seed(0)
SIZES = ((512, 256), (1024, 512), (2048, 1024), (512, 512), (1024, 1024), (2048, 2048))
if len(sys.argv) > 2:
    SIZES = ((int(sys.argv[1]), int(sys.argv[2])),)
for m, n in SIZES:
    A = torch.randn(m, n, dtype=torch.float64, device=be.device, generator=g)
    # Schmidt-like graded spectrum
    U, S, Vt = torch.linalg.svd(A, full_matrices=False)
    A = (U * torch.logspace(0, -10, n, dtype=torch.float64, device=be.device)) @ Vt
    be.svd(A)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    U1, S1, V1 = be.svd(A)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    S2 = torch.linalg.svdvals(A)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    U3, S3, V3 = torch.linalg.svd(A, full_matrices=False)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    Q, R = torch.linalg.qr(A)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    err = float((S1 - S3).abs().max() / S3.max())
    print('%dx%d  jacobi %.1f ms (%d sweeps)  cusolver svdvals %.1f ms  svd %.1f ms  qr %.2f ms  max|dS|/S0 %.1e' %
          (m, n, (t1 - t0) * 1e3, be.last_svd_sweeps, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, err))
