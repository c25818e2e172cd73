"""Synthetic effective-Hamiltonian matvec at a given site shape, for ncu captures of chain_gemm_kernel.
    ncu --set full --clock-control none --import-source on -k regex:chain_gemm -s 4 -c 4 -o gpurun_out/prof \
        python tools/profile_matvec.py --chi 1024 --kl 4 --kr 4 --nx 42 --reps 4
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnalg_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--chi', type=int, default=1024)
ap.add_argument('--d', type=int, default=2)
ap.add_argument('--kl', type=int, default=4)
ap.add_argument('--kr', type=int, default=4)
ap.add_argument('--nx', type=int, default=42)
ap.add_argument('--reps', type=int, default=4)
args = ap.parse_args()
be = ops.backend()
a = b = args.chi
d = args.d
g = torch.Generator(device=be.device).manual_seed(0)


def sym(n):
    m = torch.randn(n, n, dtype=torch.float64, device=be.device, generator=g)
    return (m + m.t()) / 2


sp = [np.array([[0.5, 0], [0, -0.5]]), np.array([[0., 1], [0, 0]]), np.array([[0., 0], [1, 0]])]
plan = be.effh_plan((a, d, b), sym(a), sym(b), None, [sym(a) for _ in range(args.kl - 1)], sp[:args.kl - 1],
                    [sym(b) for _ in range(args.kr - 1)], sp[:args.kr - 1], [sym(a) for _ in range(args.nx)],
                    [sym(b) for _ in range(args.nx)], [0.5] * args.nx)
x = torch.randn(a, d, b, dtype=torch.float64, device=be.device, generator=g)
y = torch.empty_like(x)
for _ in range(2):
    plan.matvec(x, 0.0, 1.0, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    plan.matvec(x, 0.0, 1.0, out=y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
print('matvec %.3f ms  algorithmic %.2f TFLOP/s  (a=b=%d K_L=%d K_R=%d n_x=%d)' %
      (ms, plan.flops_algorithmic / ms / 1e9, a, args.kl, args.kr, args.nx))
