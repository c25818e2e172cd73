"""One Lanczos solve on a synthetic chain-like plan (K_L = K_R = 4, n_x = 0); prints wall time per step.
    ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file out.csv python tools/profile_lanczos.py --chi 256"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnalg_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--chi', type=int, default=256)
ap.add_argument('--nx', type=int, default=0)
ap.add_argument('--solves', type=int, default=5)
args = ap.parse_args()
be = ops.backend()
a = b = args.chi
g = torch.Generator(device=be.device).manual_seed(0)


def sym(n):
    m = torch.randn(n, n, dtype=torch.float64, device=be.device, generator=g)
    return (m + m.t()) / 2


sp = [np.array([[0.5, 0], [0, -0.5]]), np.array([[0., 1], [0, 0]]), np.array([[0., 0], [1, 0]])]
plan = be.effh_plan((a, 2, b), sym(a), sym(b), None, [sym(a) for _ in range(3)], sp, [sym(b) for _ in range(3)], sp,
                    [sym(a) for _ in range(args.nx)], [sym(b) for _ in range(args.nx)], [0.5] * args.nx)
x = torch.randn(a * 2 * b, dtype=torch.float64, device=be.device, generator=g)
for _ in range(2):
    be.lanczos(plan, 1e-4, x, 1e-5, ncv=20, max_restarts=1)
torch.cuda.synchronize()
t0 = time.perf_counter()
n_mv = 0
for _ in range(args.solves):
    lam, vec, nm, resid, ok = be.lanczos(plan, 1e-4, x, 1e-5, ncv=20, max_restarts=1)
    n_mv += nm
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print('chi=%d  %d matvecs  %.1f us per Lanczos step (wall)' % (args.chi, n_mv, dt / n_mv * 1e6))
