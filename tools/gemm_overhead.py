"""Fixed cost of one chain-GEMM stage launch: time the LEFT (or RIGHT) stage alone for 1, 2, 4, 8, 16 links at a given chi
(CUDA events, back-to-back launches); slope = time per link, intercept = fixed overhead per launch."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnalg_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--chis', default='256,512,1024')
ap.add_argument('--reps', type=int, default=40)
args = ap.parse_args()
be = ops.backend()
g = torch.Generator(device=be.device).manual_seed(0)
sz = np.array([[0.5, 0], [0, -0.5]])


def sym(n):
    m = torch.randn(n, n, dtype=torch.float64, device=be.device, generator=g)
    return (m + m.t()) / 2


for chi in [int(c) for c in args.chis.split(',')]:
    a = b = chi
    x = torch.randn(a, 2, b, dtype=torch.float64, device=be.device, generator=g)
    y = torch.empty_like(x)
    for side in ('left', 'right'):
        res = []
        for nl in (1, 2, 4, 8, 16):
            mats = [sym(a) for _ in range(nl)]
            kw = dict(LS=mats, ls_ops=[sz] * nl) if side == 'left' else dict(RS=mats, rs_ops=[sz] * nl)
            plan = be.effh_plan((a, 2, b), None, None, None, **kw)
            for _ in range(3):
                plan.matvec(x, 0.0, 1.0, out=y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                plan.matvec(x, 0.0, 1.0, out=y)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / args.reps * 1e3
            fl = 2.0 * a * 2 * b * a * nl
            res.append((nl, us, fl / us / 1e6))
            plan.destroy()
        slope = (res[-1][1] - res[-2][1]) / 8
        print('chi=%d %s: ' % (chi, side) + '  '.join('%d links %.1f us (%.1f TF/s)' % r for r in res) +
              '  | per link %.1f us (floor %.1f), intercept %.1f us' % (slope, 2.0 * a * 2 * b * a / 37.09e6, res[-1][1] - 16 * slope))
