"""Two-site update mode (host logic on the oracle-backed CPU backend): the grouped two-site effective Hamiltonian against
the dense projected Hamiltonian P^T H P, and the sweep driver against dense-solve two-site DMRG and exact diagonalisation."""
import numpy as np
import pytest

from oracle import dmrg_oracle as orc
from tests.cpu_backend import CpuBackend, install as cpu_backend_install


@pytest.fixture()
def cpu_be():
    from tnalg_b200 import ops
    old = ops._backend
    be = CpuBackend()
    cpu_backend_install(be)
    yield be
    cpu_backend_install(old)


def small_para(kind, **kw):
    from tnalg_b200 import Parameters as Pm
    if kind == 'xxz':
        para = Pm.generate_parameters_dmrg('chain')
        para.update(l=8, jxy=1.0, jz=0.6, hx=0.25, hz=0.1)
    else:
        para = dict(orc.j1j2_square_para(3, 2, j1=1.0, j2=0.5))
        para.update(hx=0.0, hz=0.0)
    para.update(kw)
    if kind == 'xxz':
        para = Pm.make_consistent_parameter_dmrg(para)
    return para


@pytest.mark.parametrize('kind', ['xxz', 'j1j2'])
@pytest.mark.parametrize('p', [0, 1, 3])
def test_two_site_plan_matches_projected_dense_hamiltonian(cpu_be, kind, p):
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = small_para(kind, chi=5)
    L, d = para['l'], para['d']
    p = min(p, L - 2)
    np.random.seed(3 + p)
    A = MpsOpenBoundaryClass(L, d, para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(p)
    env = A._environments(para['index1'], para['index2'], para['coeff1'], para['coeff2'], 1e-12)
    plan = env.plan_two_site(p, A.mps)
    host = [np.asarray(t.numpy()) for t in A.mps]
    heff = orc.dense_two_site_effective_hamiltonian(host, p, para)
    a, b = host[p].shape[0], host[p + 1].shape[2]
    x = np.random.randn(a, d * d, b)
    y = plan.matvec(cpu_be.from_numpy(x), 0.0, 1.0).numpy().reshape(-1)
    ref = heff @ x.reshape(-1)
    assert np.abs(y - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    assert abs(heff - heff.T).max() < 1e-12


def test_two_site_sweep_grows_bonds_and_reaches_exact_energy(cpu_be):
    from tnalg_b200 import DMRG_anyH
    para = small_para('xxz', chi=16, sweep_time=6, dt_ob=1, break_tol=1e-13, eigs_tol=1e-14)
    np.random.seed(1)
    ob, A, info, _ = DMRG_anyH.dmrg_finite_size_two_site(para, chi_init=2)
    e0 = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert abs(ob['e_per_site'] * para['l'] - e0) < 1e-10 * abs(e0)
    assert list(A.virtual_dim) == [1, 2, 4, 8, 16, 8, 4, 2, 1]          # grown from 2 to the exact Schmidt ranks
    assert all(abs(np.linalg.norm(lm) - 1) < 1e-12 for lm in A.lm if np.size(lm))


def test_two_site_truncated_sweep_matches_dense_solve_dmrg(cpu_be):
    from tnalg_b200 import DMRG_anyH
    para = small_para('xxz', chi=4, sweep_time=30, dt_ob=1, break_tol=-1.0, eigs_tol=1e-14)
    np.random.seed(2)
    ob, A, info, _ = DMRG_anyH.dmrg_finite_size_two_site(para, chi_init=2)
    e_ref, mps_ref, lm_ref = orc.dmrg_two_site_dense(para, 4, 30, seed=5)
    e = ob['e_per_site'] * para['l']
    e0 = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert e >= e0 - 1e-12 and max(A.virtual_dim) == 4
    # both are variational chi=4 fixed points of the same algorithm; the truncated-state energies agree
    e_ref_state = orc.mps_overlap(mps_ref, mps_ref)
    psi = orc.two_site_isometry(mps_ref, 0) @ np.einsum('asb,btc->astc', mps_ref[0], mps_ref[1]).reshape(-1)
    e_ref_var = psi @ orc.dense_hamiltonian(para) @ psi / (psi @ psi)
    assert abs(e - e_ref_var) < 1e-8 * abs(e0), (e, e_ref_var, e_ref, e_ref_state)
    mid = para['l'] // 2 - 1
    # A.lm holds the Schmidt spectrum of the final (truncated) state: compare with the oracle state's
    sv = np.linalg.svd((psi / np.linalg.norm(psi)).reshape(para['d'] ** (mid + 1), -1), compute_uv=False)[:4]
    assert np.abs(np.asarray(A.lm[mid]) - sv).max() < 1e-7
    assert lm_ref[mid].size == 4


@pytest.mark.parametrize('lattice,kw', [('chain', dict(l=8, bound_cond='periodic', jxy=1, jz=0.8, hx=0.2, hz=0)),
                                        ('longRange', dict(l=8, jxy=0, jz=1, hx=0.5, hz=0, alpha=1.0)),
                                        ('square', dict(square_width=4, square_height=2))])
def test_two_site_sweep_on_other_lattices_reaches_exact_energy(cpu_be, lattice, kw):
    """long bonds (ring closure, power-law couplings, 2D snake): crossing / in-window / window-edge terms of the pair grouping"""
    from tnalg_b200 import DMRG_anyH, Parameters as Pm
    para = Pm.generate_parameters_dmrg(lattice)
    para.update(chi=16, sweep_time=8, dt_ob=1, break_tol=1e-13, eigs_tol=1e-14, **kw)
    if lattice == 'square':
        para['op'] = para['op'][:6]
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(3)
    ob, A, info, _ = DMRG_anyH.dmrg_finite_size_two_site(para, chi_init=2)
    e0 = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert abs(float(np.ravel(ob['e_per_site'])[0]) * para['l'] - e0) < 1e-10 * abs(e0)
    assert max(A.virtual_dim) == 16 and info['not_converged'] == 0
