"""Two-site update mode on the B200 (d*d = 4 physical index through the same plan / chain-GEMM kernels, Jacobi SVD
truncation): parity with the dense projected Hamiltonian, dense-solve two-site DMRG and exact diagonalisation."""
import numpy as np
import pytest

from oracle import dmrg_oracle as orc
from tests.test_two_site_cpu import small_para

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('kind', ['xxz', 'j1j2'])
@pytest.mark.parametrize('p', [0, 1, 3])
def test_two_site_plan_matches_projected_dense_hamiltonian(kind, p):
    from tnalg_b200 import ops
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    be = ops.backend()
    para = small_para(kind, chi=5)
    L, d = para['l'], para['d']
    p = min(p, L - 2)
    np.random.seed(3 + p)
    A = MpsOpenBoundaryClass(L, d, para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(p)
    env = A._environments(para['index1'], para['index2'], para['coeff1'], para['coeff2'], 1e-12)
    plan = env.plan_two_site(p, A.mps)
    host = [be.to_numpy(t) for t in A.mps]
    heff = orc.dense_two_site_effective_hamiltonian(host, p, para)
    a, b = host[p].shape[0], host[p + 1].shape[2]
    x = np.random.randn(a, d * d, b)
    y = be.to_numpy(plan.matvec(be.from_numpy(x), 0.0, 1.0)).reshape(-1)
    ref = heff @ x.reshape(-1)
    assert np.abs(y - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())


def test_two_site_sweep_reaches_exact_energy_and_grows_bonds():
    from tnalg_b200 import DMRG_anyH
    para = small_para('xxz', chi=16, sweep_time=6, dt_ob=1, break_tol=1e-13, eigs_tol=1e-14)
    np.random.seed(1)
    ob, A, info, _ = DMRG_anyH.dmrg_finite_size_two_site(para, chi_init=2)
    e0 = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert abs(float(np.ravel(ob['e_per_site'])[0]) * para['l'] - e0) < 1e-10 * abs(e0)
    assert list(A.virtual_dim) == [1, 2, 4, 8, 16, 8, 4, 2, 1]
    assert info['not_converged'] == 0


def test_two_site_truncated_sweep_matches_dense_solve_dmrg():
    from tnalg_b200 import DMRG_anyH
    para = small_para('xxz', chi=4, sweep_time=30, dt_ob=1, break_tol=-1.0, eigs_tol=1e-14)
    np.random.seed(2)
    ob, A, info, _ = DMRG_anyH.dmrg_finite_size_two_site(para, chi_init=2)
    e_ref, mps_ref, lm_ref = orc.dmrg_two_site_dense(para, 4, 30, seed=5)
    psi = orc.two_site_isometry(mps_ref, 0) @ np.einsum('asb,btc->astc', mps_ref[0], mps_ref[1]).reshape(-1)
    h = orc.dense_hamiltonian(para)
    e_var = psi @ h @ psi / (psi @ psi)
    e = float(np.ravel(ob['e_per_site'])[0]) * para['l']
    assert abs(e - e_var) < 1e-10 * abs(e_var)
    mid = para['l'] // 2 - 1
    sv = np.linalg.svd((psi / np.linalg.norm(psi)).reshape(para['d'] ** (mid + 1), -1), compute_uv=False)[:4]
    assert np.abs(np.asarray(A.lm[mid]) - sv).max() < 1e-8


def test_two_site_j1j2_4x2_truncated_is_variational():
    from tnalg_b200 import DMRG_anyH
    para = dict(orc.j1j2_square_para(4, 2, j1=1.0, j2=0.5))
    para.update(chi=8, sweep_time=6, dt_ob=1, break_tol=1e-10, eigs_tol=1e-12)
    np.random.seed(4)
    ob, A, info, _ = DMRG_anyH.dmrg_finite_size_two_site(para, chi_init=2)
    e = float(np.ravel(ob['e_per_site'])[0]) * para['l']
    e0 = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    e_ref, _, _ = orc.dmrg_two_site_dense(para, 8, 10, seed=1)     # dense-solve two-site DMRG at the same chi
    assert e >= e0 - 1e-10 and abs(e - e_ref) < 1e-10 * abs(e_ref)
    assert max(A.virtual_dim) == 8 and info['not_converged'] == 0


def test_two_site_chain20_chi32_agrees_with_one_site_sweep():
    """window sizes a = b = 32 with d*d = 4: the two-site and the one-site sweep at the same chi reach the same energy"""
    from tnalg_b200 import DMRG_anyH, Parameters as Pm
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=20, chi=32, sweep_time=6, dt_ob=1, break_tol=1e-11, eigs_tol=1e-13)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(6)
    ob2, A2, info2, _ = DMRG_anyH.dmrg_finite_size_two_site(dict(para), chi_init=4)
    np.random.seed(6)
    ob1, A1, info1, _ = DMRG_anyH.dmrg_finite_size(dict(para))
    e1, e2 = float(np.ravel(ob1['e_per_site'])[0]), float(np.ravel(ob2['e_per_site'])[0])
    assert abs(e1 - e2) < 1e-6 * abs(e1), (e1, e2)
    assert max(A2.virtual_dim) == 32 and info2['not_converged'] == 0
    assert all(abs(np.linalg.norm(lm) - 1) < 1e-10 for lm in A2.lm if np.size(lm))
    mid = para['l'] // 2 - 1
    assert np.abs(np.asarray(A1.lm[mid])[:8] - np.asarray(A2.lm[mid])[:8]).max() < 1e-4
