"""Exact-diagonalisation cross-checker on the device -- the part of library/EDspinClass.py:6-77 (EDbasic) and
algorithms/ExactDiagonalizationAlgo.py:12-24 (exact_ground_state) that serves as a known-answer check of the DMRG path.

EDbasic keeps the full d^L state on the GPU; project_all_hamilt is one gather kernel (tn_ed_apply) and the ground state runs
the same device-resident Lanczos that solves the DMRG local problems (tn_ed_ground_state).  Uniform physical dimensions only
(the DMRG path never mixes them); L is limited by d^L <= 2^31.
"""
import numpy as np

from . import ops as _ops


class EDbasic:
    def __init__(self, dims, state_type='pure', ini='random', operators=None):
        dims = [int(x) for x in dims]
        if len(set(dims)) != 1:
            raise ValueError('EDbasic (device): all physical dimensions must be equal')
        if state_type != 'pure':
            raise ValueError('EDbasic (device): only pure states')
        self.dims, self.l, self.d = dims, len(dims), dims[0]
        self.dim_tot = int(np.prod(dims))
        self.state_type = state_type
        self._be = _ops.backend()
        if isinstance(ini, np.ndarray):
            v = np.real(ini).reshape(-1).astype(float)
        elif ini == 'random':
            v = np.random.randn(self.dim_tot)
            v /= np.linalg.norm(v)
        else:
            v = np.ones(self.dim_tot) / self.dim_tot ** 0.5
        self.v = self._be.from_numpy(v)
        self.operators = operators

    def project_all_hamilt(self, v0, hamilts, tau, couplings):
        """v0 - tau * sum_n hamilts[c_n] on sites (p1_n, p2_n)  (EDspinClass.py:69-77); numpy in -> numpy out, tensor in -> tensor out"""
        be = self._be
        dev = hasattr(v0, 'data_ptr')
        x = v0 if dev else be.from_numpy(np.real(np.asarray(v0)).reshape(-1))
        y = be.ed_apply(x, self.l, self.d, couplings, hamilts, 1.0, -float(tau))
        return y if dev else be.to_numpy(y)

    def ground_state(self, hamilts, couplings, tau=1e-4, tol=1e-12):
        """the eigs(heff, k=1, which='LM') of exact_ground_state: returns (E0, state as numpy vector); self.v is updated"""
        lam, vec, n_mv, resid, ok = self._be.ed_ground_state(self.v, self.l, self.d, couplings, hamilts, tau=tau, tol=tol)
        self.v = vec
        self.last = {'n_matvec': n_mv, 'residual': resid, 'converged': ok}
        return (1.0 - lam) / tau, self._be.to_numpy(vec)

    def reduced_matrix(self, bonds1):
        """reduced density matrix of the sites `bonds1` (EDspinClass.py:79-93), on the host (a check helper)"""
        v = self._be.to_numpy(self.v).reshape(self.dims)
        bonds2 = [b for b in range(self.l) if b not in bonds1]
        mat = v.transpose(list(bonds1) + bonds2).reshape(int(np.prod([self.dims[b] for b in bonds1])), -1)
        return mat.conj() @ mat.T

    def observe_operator(self, op, bonds):
        mat = self.reduced_matrix(bonds)
        return np.trace(mat.dot(op)) / np.trace(mat)


def exact_ground_state(para):
    """ground state of sum_bonds h(p1, p2) for a para dict carrying `l, d, tau, positions_h2` and the two-site matrix
    para['hamilt'] (d^2, d^2): returns (EDbasic, ob) with ob['e_eig'], ob['e_site'] like ExactDiagonalizationAlgo.py:12-36"""
    pos = np.asarray(para['positions_h2'], dtype=int).reshape(-1, 2)
    couplings = np.hstack([pos, np.zeros((pos.shape[0], 1), dtype=int)])
    a = EDbasic([para['d']] * para['l'])
    e0, _ = a.ground_state([para['hamilt']], couplings, tau=para.get('tau', 1e-4), tol=para.get('eigs_tol', 1e-12))
    ob = {'e_eig': e0, 'eb': [float(np.real(a.observe_operator(para['hamilt'], [int(p[0]), int(p[1])]))) for p in pos]}
    ob['e_site'] = sum(ob['eb']) / para['l']
    return a, ob
