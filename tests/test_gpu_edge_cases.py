"""Edge cases of the path through the real kernels (tiny and ragged shapes, product states, chi larger than the exact
bond dimension, field-only and Ising term lists, spin-1): exact cases against ED, truncated ones variational."""
import numpy as np
import pytest

from oracle import dmrg_oracle as orc

pytestmark = pytest.mark.gpu

CASES = [
    ('L2', dict(l=2, chi=4), True),
    ('L3_product_state', dict(l=3, chi=1), False),
    ('L4_chi2', dict(l=4, chi=2), False),
    ('L6_chi_above_exact', dict(l=6, chi=64), True),
    ('L5_field_only', dict(l=5, chi=4, jxy=0, jz=0, hx=0.5, hz=0.2), True),
    ('L6_ising', dict(l=6, chi=8, jxy=0, jz=1, hx=0.7), True),
    ('L6_spin_one_exact', dict(l=6, chi=27, spin='one'), True),
    ('L6_spin_one_truncated', dict(l=6, chi=9, spin='one'), False),
]


def _para(kw):
    from tnalg_b200 import Parameters as Pm
    para = Pm.generate_parameters_dmrg('chain')
    para.update(eigs_tol=1e-12, break_tol=1e-13, **kw)
    return Pm.make_consistent_parameter_dmrg(para)


@pytest.mark.parametrize('two_site', [False, True])
@pytest.mark.parametrize('name,kw,exact', CASES, ids=[c[0] for c in CASES])
def test_edge_case_energy(name, kw, exact, two_site):
    from tnalg_b200 import DMRG_anyH
    para = _para(kw)
    np.random.seed(0)
    if two_site:     # d*d = 9 (spin-1) runs too: site operators beyond TN_MAX_LOADPATH_DIM are pre-applied element-wise
        ob, A, info, _ = DMRG_anyH.dmrg_finite_size_two_site(para, chi_init=min(2, para['chi']))
    else:
        ob, A, info, _ = DMRG_anyH.dmrg_finite_size(para)
    e = float(np.ravel(ob['e_per_site'])[0]) * para['l']
    e0 = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    if exact:
        assert abs(e - e0) <= 1e-10 * abs(e0), (name, e, e0)
    else:
        assert e0 - 1e-10 <= e < 0.0, (name, e, e0)
    assert max(A.virtual_dim) <= para['chi'] and info['not_converged'] == 0
    for lm in A.lm:
        if np.size(lm):
            assert abs(np.linalg.norm(lm) - 1) < 1e-10


def test_zero_coefficient_terms_are_dropped():
    """terms with |c| <= tol or a zero operator are filtered before the plan is built (MPSClass.py:684-733 uses tol = eigs_tol)"""
    from tnalg_b200 import DMRG_anyH
    para = _para(dict(l=6, chi=8))
    n2 = para['index2'].shape[0]
    extra = np.array([[0, 5, 3, 3], [1, 4, 1, 1]])
    p2 = dict(para)
    p2['index2'] = np.vstack([para['index2'], extra])
    p2['coeff2'] = np.concatenate([np.ravel(para['coeff2']), [0.0, 1e-14]]).reshape(-1, 1)
    p2['nh'] = n2 + 2
    np.random.seed(0)
    ob1, A1, _, _ = DMRG_anyH.dmrg_finite_size(dict(para))
    np.random.seed(0)
    ob2, A2, _, _ = DMRG_anyH.dmrg_finite_size(p2)
    e1, e2 = float(np.ravel(ob1['e_per_site'])[0]), float(np.ravel(ob2['e_per_site'])[0])
    assert abs(e1 - e2) < 1e-13, (e1, e2)      # the 1e-14 term still enters the reported bond-energy sum
    # ... but not the optimisation: same state (bit-identical on the CPU stand-in; the GPU matvec adds partial tiles with
    # FP64 atomics, so two runs agree to round-off only)
    assert all(np.abs(np.asarray(x) - np.asarray(y)).max() < 1e-9 for x, y in zip(A1.lm, A2.lm))
