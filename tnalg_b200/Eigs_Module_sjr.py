"""eigs_fh with the reference's signature (Eigs_Module_sjr.py:18-58) for effective-Hamiltonian plans.

The reference routine is a host Lanczos with explicit Gram-Schmidt that nothing calls and that
fails on its own restart path (np.random.randn(w.shape), :77,98,101).  This wrapper keeps the call shape
    lm, v, info = eigs_fh(lin_map, d, n=1, k=-1, v0, tol, max_it, which='lm')
but `lin_map` must be a tnalg_b200 EffHPlan (the matvec lives on the device) and only n = 1 is supported: the
extreme eigenpair is computed by the device-resident thick-restart Lanczos (tn_lanczos_lm1).
which: 'sa' -> lowest eigenvalue of H_eff, 'la' -> highest, 'lm' -> largest magnitude.
"""
import numpy as np

from . import ops as _ops
from .ops import EffHPlan


def eigs_fh(lin_map, d, n=1, k=-1, v0=np.zeros(0), tol=1e-15, max_it=1000, which='lm'):
    if not isinstance(lin_map, EffHPlan):
        raise TypeError('eigs_fh: lin_map must be a tnalg_b200.ops.EffHPlan (device matvec); host callables are not '
                        'supported -- there is no CPU path')
    if n != 1:
        raise NotImplementedError('eigs_fh: only the extreme eigenpair (n=1) is computed on the device')
    be = _ops.backend()
    dim = int(np.prod(lin_map.shape))
    if d != dim:
        raise ValueError('eigs_fh: d=%d does not match the plan dimension %d' % (d, dim))
    if hasattr(v0, 'data_ptr'):
        v = v0.reshape(-1)
    elif np.asarray(v0).size == 0:
        v = be.from_numpy(np.random.randn(dim))
    else:
        v = be.from_numpy(np.asarray(v0, dtype=float).reshape(-1))
    ncv = 20 if k < 0 else max(k, 2)
    which = which.lower()
    # (1 - tau*H) dominant eigenpair: tau > 0 small picks the lowest eigenvalue of H, tau < 0 the highest;
    # |tau| large makes |1 - tau*theta| ~ |theta| (largest magnitude)
    tau = {'sa': 1e-4, 'la': -1e-4}.get(which, 1e8)
    lam, vec, n_mv, resid, ok = be.lanczos(lin_map, tau, v, tol * abs(tau) if which in ('sa', 'la') else tol, ncv=ncv,
                                           max_restarts=max(int(max_it), 1))
    theta = (1.0 - lam) / tau
    info = {'it_time': n_mv, 'error': np.array([resid]), 'converged': ok}
    return np.array([theta]), vec.reshape(-1, 1), info
