"""Time tn_qr_householder against cuSOLVER (torch.linalg.qr, the checked baseline) on the gauge-move shapes of the BASELINE
configurations, and report orthogonality / reconstruction errors.  Run on the GPU box:  python tools/profile_qr.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnalg_b200 import ops  # noqa: E402


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    be = ops.backend()
    rng = np.random.RandomState(0)
    print('| m x n | ours ms | cuSOLVER ms | ratio | orth err | recon err |')
    print('|---|---|---|---|---|---|')
    for m, n in [(128, 64), (512, 256), (1024, 512), (2048, 1024), (4096, 2048), (512, 512), (1024, 1024), (2048, 2048), (256, 512)]:
        a = rng.randn(m, n)
        u, s, vt = np.linalg.svd(a, full_matrices=False)
        a = (u * np.logspace(0, -10, s.size)) @ vt
        A = be.from_numpy(a)
        t_own = timeit(lambda: be.qr(A))
        t_lib = timeit(lambda: torch.linalg.qr(A, mode='reduced'))
        Q, R = be.qr(A)
        k = min(m, n)
        orth = float((Q.t() @ Q - torch.eye(k, dtype=torch.float64, device=Q.device)).abs().max())
        rec = float((Q @ R - A).abs().max() / A.abs().max())
        print('| %d x %d | %.3f | %.3f | %.2f | %.1e | %.1e |' % (m, n, t_own, t_lib, t_lib / t_own, orth, rec))


if __name__ == '__main__':
    main()
