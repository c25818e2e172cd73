"""Phase breakdown of one DMRG sweep (synchronising timers; for diagnosis only, not a benchmark)."""
import argparse
import collections
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tnalg_b200 import MPSClass, envs, ops  # noqa: E402
from tnalg_b200.DMRG_anyH import sweep_once  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='j1j2_6x6_chi1024')
ap.add_argument('--chi', type=int, default=0)
args = ap.parse_args()
para = bench.build_para(bench.WORKLOADS[args.workload], args.chi)
be = ops.backend()
acc = collections.Counter()
cnt = collections.Counter()


def timed(name, fn):
    def wrapper(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        acc[name] += time.perf_counter() - t0
        cnt[name] += 1
        return r
    return wrapper


Mps = MPSClass.MpsOpenBoundaryClass
Mps.correct_orthogonal_center = timed('gauge move (QR + absorb)', Mps.correct_orthogonal_center)
envs.EnvCache.ensure = timed('env.ensure (tn_env_update)', envs.EnvCache.ensure)
envs.EnvCache.groups = timed('env.groups (lincombs, python)', envs.EnvCache.groups)
be.effh_plan = timed('effh_plan create', be.effh_plan)
be.lanczos = timed('lanczos', be.lanczos)
be.qr = timed('  of which qr', be.qr)
be.mode_product = timed('  of which mode_product', be.mode_product)
np.random.seed(0)
A = Mps(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
A.correct_orthogonal_center(0)
sweep_once(A, para)
acc.clear()
cnt.clear()
torch.cuda.synchronize()
t0 = time.perf_counter()
sweep_once(A, para)
torch.cuda.synchronize()
total = time.perf_counter() - t0
print('sweep total %.1f ms' % (total * 1e3))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print('%-34s %8.1f ms  %5.1f %%  (%d calls)' % (k, v * 1e3, 100 * v / total, cnt[k]))
