"""reproduce the NaN seen when the chi=256 chain workload runs after the chi=1024 lattice workload in one process"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tnalg_b200 import ops  # noqa: E402
from tnalg_b200.DMRG_anyH import sweep_once  # noqa: E402
from tnalg_b200.MPSClass import MpsOpenBoundaryClass  # noqa: E402

be = ops.backend()
first = sys.argv[1] if len(sys.argv) > 1 else 'j1j2_6x6_chi1024'


def finite(t):
    return bool(torch.isfinite(t).all())


def run(name, sweeps, check):
    para = bench.build_para(bench.WORKLOADS[name])
    np.random.seed(0)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(para['ob_position'])
    if not check:
        for _ in range(sweeps):
            sweep_once(A, para)
        return
    # instrumented sweep
    qr_real, env_real, lz_real = be.qr_tensor, be.env_update, be.lanczos

    def qr_chk(T, l2r):
        Q, R = qr_real(T, l2r)
        if not (finite(Q) and finite(R)):
            torch.save({'T': T.cpu(), 'l2r': l2r}, 'gpurun_out/r2f/bad_qr.pt')
            raise RuntimeError('QR produced non-finite values: T %s finite=%s l2r=%s |T|max=%g' % (tuple(T.shape), finite(T), l2r, float(T.abs().max())))
        return Q, R

    def env_chk(direction, T, outputs):
        res = env_real(direction, T, outputs)
        for r in res:
            if not finite(r):
                raise RuntimeError('env_update produced non-finite values: T %s finite=%s' % (tuple(T.shape), finite(T)))
        return res

    def lz_chk(plan, tau, v0, tol, **kw):
        if not finite(v0):
            raise RuntimeError('Lanczos start vector is not finite, shape %s' % (plan.shape,))
        y = plan.matvec(v0.reshape(plan.shape))
        if not finite(y):
            raise RuntimeError('matvec output is not finite, shape %s' % (plan.shape,))
        return lz_real(plan, tau, v0, tol, **kw)
    be.qr_tensor, be.env_update, be.lanczos = qr_chk, env_chk, lz_chk
    try:
        for s in range(sweeps):
            for n in bench.__dict__.get('sweep_order', None) or __import__('tnalg_b200.DMRG_anyH', fromlist=['sweep_order']).sweep_order(para['l'], para['ob_position']):
                try:
                    A.update_tensor_eigs(n, para['index1'], para['index2'], para['coeff1'], para['coeff2'], para['tau'], para['is_real'], tol=para['eigs_tol'])
                except Exception as e:
                    print('FAILED at sweep %d site %d: %s' % (s, n, e))
                    raise
    finally:
        be.qr_tensor, be.env_update, be.lanczos = qr_real, env_real, lz_real
    print('ok', name)


os.makedirs('gpurun_out/r2f', exist_ok=True)
if first != 'none':
    run(first, 1, False)
    torch.cuda.empty_cache()
run('heis_chain100_chi256', 2, True)
