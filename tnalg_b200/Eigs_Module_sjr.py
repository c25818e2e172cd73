"""eigs_fh with the reference's signature (Eigs_Module_sjr.py:18-58), solved by the device-resident Lanczos.

    lm, v, info = eigs_fh(lin_map, d, n=6, k=-1, v0=np.zeros(0), tol=1e-15, max_it=1, which='lm')

The reference routine is a host Lanczos with explicit Gram-Schmidt that nothing calls and that fails on
its own restart path (np.random.randn(w.shape), :77,98,101); its call shape and return values are mirrored, not its bugs:
  * lin_map  -- an EffHPlan (the effective Hamiltonian of a site, applied by the chain-GEMM kernels), or any callable
                v -> H v.  A callable is handed a float64 CUDA tensor of shape (d, 1) and may return a CUDA tensor or a numpy
                array / anything np.asarray accepts (a host callable is supported: its result is copied to the device; the
                Krylov basis, re-orthogonalisation, projected eigenproblem and convergence test stay on the GPU);
  * d, n, k  -- dimension, number of eigenpairs, Krylov dimension per restart cycle (k < 0: min(d, 20), the ARPACK default
                the live path of the reference uses; the reference's k = d "full tridiagonalisation" is not copied);
  * which    -- 'sa' / 'la' / 'lm' / 'sm' are taken on the spectrum of lin_map as in handle_which (:126-139);
  * returns  -- lm (n,), v (d, n) numpy arrays and info = {'it_time', 'error'} like the reference
                (error[i] = residual estimate of pair i), plus 'converged' and 'n_matvec'.
Eigenpairs beyond the first are computed by deflation (each solve is kept orthogonal to the converged vectors), so they are
as accurate as the first -- the reference warns that its higher pairs "are computed badly" (:13-15).
"""
import numpy as np

from . import ops as _ops
from .ops import EffHPlan


def _tau_for(which, scale):
    """the solver returns the dominant eigenpair of 1 - tau*H: tau small positive -> lowest eigenvalue of H, small negative ->
    highest; a huge |tau| makes |1 - tau*theta| ~ |tau*theta| (largest magnitude)"""
    which = which.lower()
    if which == 'sa':
        return 1e-4 / max(scale, 1e-300)
    if which == 'la':
        return -1e-4 / max(scale, 1e-300)
    if which == 'sm':
        raise NotImplementedError("eigs_fh: which='sm' needs a shift-invert operator; not provided")
    return 1e8 / max(scale, 1e-300)


def eigs_fh(lin_map, d, n=6, k=-1, v0=np.zeros(0), tol=1e-15, max_it=1, which='lm'):
    import torch
    be = _ops.backend()
    d = int(d)
    n = int(min(n, d))
    if isinstance(lin_map, EffHPlan):
        dim = int(np.prod(lin_map.shape))
        if d != dim:
            raise ValueError('eigs_fh: d=%d does not match the plan dimension %d' % (d, dim))
        shape = lin_map.shape

        def apply(x, y):
            lin_map.matvec(x.reshape(shape), 0.0, 1.0, out=y.reshape(shape))
    elif callable(lin_map):
        def apply(x, y):
            r = lin_map(x.reshape(d, 1))
            if not hasattr(r, 'data_ptr'):
                r = be.from_numpy(np.real(np.asarray(r)).reshape(-1))
            y.copy_(r.reshape(-1))
    else:
        raise TypeError('eigs_fh: lin_map must be a callable or a tnalg_b200.ops.EffHPlan')
    if hasattr(v0, 'data_ptr'):
        v = v0.reshape(-1).clone()
    elif np.asarray(v0).size == 0:
        v = be.from_numpy(np.random.randn(d))          # set_initial_v (:118-122)
    else:
        v = be.from_numpy(np.real(np.asarray(v0, dtype=complex)).reshape(-1))
    ncv = min(d, 20) if k < 0 else min(max(int(k), n + 1, 2), d)
    # a scale for the shift: |<v|H|v>| of the start vector (tau only has to be small against 1/|spectrum|)
    probe = torch.empty_like(v)
    apply(v, probe)
    scale = be.norm(probe) / max(be.norm(v), 1e-300)
    tau = _tau_for(which, scale)
    tol_solver = max(float(tol), 5e-14) * (abs(tau) * scale if which.lower() in ('sa', 'la') else 1.0)   # floor: FP64 rounding of the residual
    restarts = max(int(max_it), 1) if max_it > 1 else 1000
    lms, vecs, errs, n_mv, ok_all = [], [], [], 0, True
    locked = None
    for i in range(n):
        start = v if i == 0 else be.from_numpy(np.random.RandomState(1234 + i).randn(d))
        if isinstance(lin_map, EffHPlan) and i == 0:
            lam, vec, mv, resid, ok = be.lanczos(lin_map, tau, start, tol_solver, ncv=ncv, max_restarts=restarts)
        else:
            lam, vec, mv, resid, ok = be.lanczos_generic(apply, d, tau, start, tol_solver, ncv=ncv, max_restarts=restarts, locked=locked)
        lms.append((1.0 - lam) / tau)
        vecs.append(vec.reshape(1, -1))
        errs.append(resid)
        n_mv += mv
        ok_all = ok_all and ok
        locked = torch.cat(vecs, 0).contiguous()
    info = {'it_time': n_mv, 'error': np.array(errs).reshape(1, -1), 'converged': ok_all, 'n_matvec': n_mv}
    return np.array(lms), be.to_numpy(locked).T.copy(), info
