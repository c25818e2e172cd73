"""Time one two-site sweep (Lanczos on theta[a, d*d, b] + Jacobi-SVD truncation) on a Heisenberg chain.
usage: python tools/profile_two_site.py [L] [chi] [n_sweeps]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from tnalg_b200 import DMRG_anyH, Parameters as Pm  # noqa: E402
from tnalg_b200.MPSClass import MpsOpenBoundaryClass  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    n_sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=L, chi=chi, eigs_tol=1e-8)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(0)
    A = MpsOpenBoundaryClass(L, para['d'], min(chi, 8), operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(0)
    A.timing = True
    be = A._be
    svd_stat = {'ms': 0.0, 'calls': 0, 'sweeps': 0}
    raw_svd = be.svd

    def timed_svd(*a, **k):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = raw_svd(*a, **k)
        torch.cuda.synchronize()
        svd_stat['ms'] += (time.perf_counter() - t) * 1e3
        svd_stat['calls'] += 1
        svd_stat['sweeps'] += getattr(be, 'last_svd_sweeps', 0)
        return r
    if torch.cuda.is_available():
        be.svd = timed_svd
    out = []
    for s in range(n_sweeps):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t0 = time.time()
        mv0 = A.stats['n_matvec']
        svd_stat.update(ms=0.0, calls=0, sweeps=0)
        A._events = []
        DMRG_anyH.sweep_once_two_site(A, para)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dt = time.time() - t0
        solver = A.solver_time_ms() if torch.cuda.is_available() else float('nan')
        ob = {}
        DMRG_anyH.observe(A, para, ob)
        out.append({'sweep': s, 'ms': dt * 1e3, 'lanczos_ms': solver, 'n_matvec': A.stats['n_matvec'] - mv0, 'svd_ms': svd_stat['ms'], 'svd_calls': svd_stat['calls'],
                    'jacobi_sweeps_mean': svd_stat['sweeps'] / max(1, svd_stat['calls']),
                    'chi_max': int(max(A.virtual_dim)), 'e_per_site': float(np.ravel(ob['e_per_site'])[0])})
        print(json.dumps(out[-1]), flush=True)


if __name__ == '__main__':
    main()
