"""Diagnose a non-converging tn_svd_jacobi call inside a two-site sweep: re-run the failing input on every kernel variant."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from tnalg_b200 import DMRG_anyH, Parameters as Pm, ops  # noqa: E402
from tnalg_b200._lib import TnError  # noqa: E402
from tnalg_b200.MPSClass import MpsOpenBoundaryClass  # noqa: E402

L, chi = int(sys.argv[1]), int(sys.argv[2])
para = Pm.generate_parameters_dmrg('chain')
para.update(l=L, chi=chi, eigs_tol=1e-8)
para = Pm.make_consistent_parameter_dmrg(para)
np.random.seed(0)
A = MpsOpenBoundaryClass(L, para['d'], min(chi, 8), operators=para['op'], is_save_op=True, eig_way=1)
A.correct_orthogonal_center(0)
be = A._be
raw = be._jacobi
fails = []


def guarded(X, k_keep=None):
    try:
        return raw(X, k_keep)
    except TnError as e:
        m, n = X.shape
        print('FAIL shape', (m, n), 'k_keep', k_keep, str(e), flush=True)
        s_ref = torch.linalg.svdvals(X)
        print('  spectrum: max %.3e  [k_keep-1] %.3e  min %.3e  #>1e-14*max %d' % (
            float(s_ref[0]), float(s_ref[(k_keep or min(m, n)) - 1]), float(s_ref[-1]), int((s_ref > 1e-14 * s_ref[0]).sum())), flush=True)
        print('  row norms of the input: max %.3e min %.3e #zero rows %d' % (
            float(X.norm(dim=1).max()), float(X.norm(dim=1).min()), int((X.norm(dim=1) == 0).sum())), flush=True)
        for name, env in (('registers/default', {}), ('smem', {'TNALG_SVD_SMEM': '1'}), ('unblocked', {'TNALG_SVD_UNBLOCKED': '1'})):
            for kk in (k_keep, None):
                os.environ.pop('TNALG_SVD_SMEM', None)
                os.environ.pop('TNALG_SVD_UNBLOCKED', None)
                os.environ.update(env)
                try:
                    raw(X, kk)
                    print('  %-18s k_keep=%s: converged in %d sweeps' % (name, kk, be.last_svd_sweeps), flush=True)
                except TnError as e2:
                    print('  %-18s k_keep=%s: %s' % (name, kk, e2), flush=True)
        os.environ.pop('TNALG_SVD_SMEM', None)
        os.environ.pop('TNALG_SVD_UNBLOCKED', None)
        torch.save(X.cpu(), 'gpurun_out/svd_noconv_input.pt')
        fails.append((m, n))
        raise


be._jacobi = guarded
try:
    for s in range(3):
        DMRG_anyH.sweep_once_two_site(A, para)
        print('sweep', s, 'ok, chi_max', int(max(A.virtual_dim)), flush=True)
except TnError:
    print('stopped after first failure')
