"""Host logic of the product (term algebra, environment cache, sweep driver, observables, pickling) exercised on
the CPU with the oracle-backed stand-in backend of tests/cpu_backend.py, against vectors produced by the
unmodified reference (tests/golden/*.npz)."""
import io
import pickle

import numpy as np
import pytest

from tests.cpu_backend import CpuBackend, install as cpu_backend_install


@pytest.fixture()
def cpu_be():
    from tnalg_b200 import ops
    old = ops._backend
    be = CpuBackend()
    cpu_backend_install(be)
    yield be
    cpu_backend_install(old)


def para_from_golden(g, **kw):
    from tnalg_b200 import Parameters as Pm
    para = dict(Pm.common_parameters_dmrg())
    ops = [np.real(o) if np.abs(np.imag(o)).max() == 0 else o for o in g['op']]
    para.update(lattice='arbitrary', spin='one' if int(g['d']) == 3 else 'half', op=ops, index1=g['index1'], coeff1=g['coeff1'], index2=g['index2'],
                coeff2=g['coeff2'], chi=int(g['chi']), tau=float(g['tau']), eigs_tol=float(g['eigs_tol']),
                break_tol=float(g['break_tol']), hx=float(g['hx']), hz=float(g['hz']))
    para.update(kw)
    para = Pm.make_consistent_parameter_dmrg(para)
    if 'positions_h2' in g:
        para['positions_h2'] = np.asarray(g['positions_h2'])   # bond order of the generator that made the golden ('square' differs)
    return para


def test_parameters_match_reference_generators(golden):
    from tnalg_b200 import Parameters as Pm
    g = golden('e2e_chain12')
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=12, chi=16)
    para = Pm.make_consistent_parameter_dmrg(para)
    for k in ('index1', 'index2', 'coeff1', 'coeff2', 'positions_h2'):
        assert np.array_equal(np.asarray(para[k]), g[k]), k
    assert len(para['op']) == 7 and all(np.allclose(a, b) for a, b in zip(para['op'], g['op']))
    assert para['data_exp'] == 'chainN12_j(1,1)_h(0,0)_chi16open' and para['d'] == 2 and para['nh'] == 33
    g = golden('e2e_xxz10')
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=10, chi=12, jxy=1, jz=0.5, hx=0.3, hz=0)
    para = Pm.make_consistent_parameter_dmrg(para)
    for k in ('index1', 'index2', 'coeff1', 'coeff2'):
        assert np.array_equal(np.asarray(para[k]), g[k]), k
    assert np.allclose(para['op'][6], g['op'][6])
    sq = Pm.generate_parameters_dmrg('square')
    assert sq['l'] == 16 and sq['index2'].shape == (72, 4) and sq['positions_h2'].shape == (24, 2)


@pytest.mark.parametrize('p', [0, 2, 4, 5, 8])
def test_plan_groups_and_matvec_vs_reference(golden, cpu_be, p):
    """the summed complementary blocks give the reference's opt_env groups and the reference's handle output"""
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    g = golden('percall_j1j2')
    L, d, chi = int(g['l']), int(g['d']), int(g['chi'])
    ops = [np.real(o) for o in g['op']]
    A = MpsOpenBoundaryClass(L, d, chi, operators=ops, is_save_op=True, eig_way=1)
    for n in range(L):
        A.mps[n] = g['p%d_mps_%d' % (p, n)]
    A.center = p
    env = A._environments(g['index1'], g['index2'], g['coeff1'], g['coeff2'], 1e-12)
    # group structure == reference key set
    kl, kr, nx = env.terms.counts(p)
    keys = [str(k) for k in g['p%d_keys' % p]]
    ref_left = sum(1 for k in keys if k.startswith('1_') and k.endswith('_0'))
    ref_right = sum(1 for k in keys if k.startswith('0_') and k.endswith('_1'))
    assert (kl, kr, nx) == (ref_left, ref_right, int(g['p%d_ncross' % p]))
    plan = A.effective_hamiltonian_plan(p, g['index1'], g['index2'], g['coeff1'], g['coeff2'], tol=1e-12)
    x = cpu_be.from_numpy(g['p%d_x' % p].reshape(tuple(g['p%d_shape' % p])))
    y = plan.matvec(x, 1.0, -float(g['tau'])).numpy().reshape(-1)
    assert np.abs(y - g['p%d_y' % p]).max() < 1e-13 * max(1.0, np.abs(g['p%d_y' % p]).max())
    hx = plan.matvec(x, 0.0, 1.0).numpy().reshape(-1)
    assert np.abs(hx - g['p%d_heff' % p] @ g['p%d_x' % p]).max() < 1e-12


def test_observables_vs_reference(golden, cpu_be):
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    g = golden('percall_j1j2')
    L, d, chi = int(g['l']), int(g['d']), int(g['chi'])
    A = MpsOpenBoundaryClass(L, d, chi, operators=[np.real(o) for o in g['op']])
    for n in range(L):
        A.mps[n] = g['ob_mps_%d' % n]
    A.center = 4
    assert np.abs(A.observe_magnetization(1) - g['ob_mx']).max() < 1e-13
    assert np.abs(A.observe_magnetization(3) - g['ob_mz']).max() < 1e-13
    assert np.abs(A.observe_bond_energy(g['index2'], g['coeff2']) - g['ob_eb_full']).max() < 1e-13
    assert np.abs(A.observe_correlators_from_middle(3, 3) - g['ob_corr_z']).max() < 1e-13
    assert np.abs(A.observe_correlators_from_middle(1, 1) - g['ob_corr_x']).max() < 1e-13
    assert abs(A.norm_mps() - float(g['ob_norm'])) < 1e-13
    assert abs(A.observation_s1((3, 2)) - g['ob_mz'][2, 0]) < 1e-13
    # centre at either end: same numbers (gauge invariance of the observable pass)
    for c in (0, L - 1):
        A.correct_orthogonal_center(c)
        assert np.abs(A.observe_magnetization(3) - g['ob_mz']).max() < 1e-12
        assert np.abs(A.observe_bond_energy(g['index2'], g['coeff2']) - g['ob_eb_full']).max() < 1e-12
        # the one-pass form the sweep driver uses (shared density chain, S+S- / S-S+ contracted once)
        eb, (mx, mz) = A.observe_bond_energy_and_magnetization(g['index2'], g['coeff2'], (1, 3))
        assert np.abs(eb - g['ob_eb_full']).max() < 1e-12 and np.abs(mx - g['ob_mx']).max() < 1e-12 and np.abs(mz - g['ob_mz']).max() < 1e-12


@pytest.mark.parametrize('case', ['e2e_chain12', 'e2e_xxz10', 'e2e_j1j2_4x2', 'e2e_spin1_chain8', 'e2e_longrange8', 'e2e_square3x2', 'e2e_periodic8'])
def test_end_to_end_vs_reference(golden, cpu_be, case):
    """dmrg_finite_size through the product's host code: converged energies / spectrum rel 1e-10, observables 1e-8"""
    from tnalg_b200.DMRG_anyH import dmrg_finite_size
    g = golden(case)
    para = para_from_golden(g)
    np.random.seed(int(g['seed']))
    ob, A, info, para = dmrg_finite_size(para)
    assert abs(ob['e_per_site'][0] - g['e_per_site'][0]) <= 1e-10 * abs(g['e_per_site'][0])
    for k in ('eb_full', 'eb', 'mx', 'mz', 'corr_x', 'corr_z'):
        assert np.abs(np.asarray(ob[k]).reshape(-1) - g[k].reshape(-1)).max() < 1e-8, k
    assert np.abs(A.ent - g['ent']).max() < 1e-8
    for n in range(para['l'] - 1):
        ref = g['lm_%d' % n]
        assert np.abs(A.lm[n] - ref).max() <= 1e-10 * ref.max() + 1e-12, n
    assert np.array_equal(A.virtual_dim, g['virtual_dim'])
    assert info['not_converged'] == 0
    # one batched bond move per local update instead of the reference's per-term chains
    assert set(ob.keys()) == {'eb_full', 'mx', 'mz', 'e_per_site', 'eb', 'corr_x', 'corr_z'}
    assert {'convergence', 't_cost'} <= set(info.keys())


def test_pr_layout_roundtrip(golden, cpu_be, tmp_path):
    """`.pr` = pickle of {name: obj}; the MPS object pickles to the reference attribute set with numpy tensors"""
    from tnalg_b200 import BasicFunctionsSJR as Bf
    from tnalg_b200.DMRG_anyH import dmrg_finite_size
    g = golden('e2e_chain12')
    para = para_from_golden(g, sweep_time=2, dt_ob=2)
    np.random.seed(0)
    ob, A, info, para = dmrg_finite_size(para)
    Bf.save_pr(str(tmp_path), 'x.pr', (ob, A, info, para), ('ob', 'A', 'info', 'para'))
    data = Bf.load_pr(str(tmp_path / 'x.pr'))
    assert set(data.keys()) == {'ob', 'A', 'info', 'para'}
    B = data['A']
    ref_attrs = set(str(k) for k in g['attrs'])                       # today's reference class (MPSClass.py:53-106)
    old_attrs = set(str(k) for k in golden('pr_fixtures')['chi16_attrs'])  # the 2018 fixtures hold a subset
    assert set(B.__dict__.keys()) == ref_attrs and old_attrs <= ref_attrs
    assert all(isinstance(t, np.ndarray) and t.dtype == np.float64 for t in B.mps)
    assert [t.shape for t in B.mps] == [(A.virtual_dim[n], 2, A.virtual_dim[n + 1]) for n in range(12)]
    ob2 = Bf.load_pr(str(tmp_path / 'x.pr'), 'ob')
    assert np.array_equal(ob2['mz'], ob['mz'])
    assert Bf.load_pr(str(tmp_path / 'missing.pr')) is False
    # a revived object keeps working (tensors go back to the device lazily)
    assert abs(B.norm_mps() - 1.0) < 1e-12
    assert np.abs(B.observe_magnetization(3) - ob['mz']).max() < 1e-12


def test_cache_invalidation_on_external_gauge_change(golden, cpu_be):
    """re-gauging between updates (what breaks the reference's corr_*) must not leave stale blocks behind"""
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    g = golden('percall_j1j2')
    L, d, chi = int(g['l']), int(g['d']), int(g['chi'])
    A = MpsOpenBoundaryClass(L, d, chi, operators=[np.real(o) for o in g['op']])
    for n in range(L):
        A.mps[n] = g['mps_%d' % n]
    A.center = int(g['center0'])
    args = (g['index1'], g['index2'], g['coeff1'], g['coeff2'])
    A.correct_orthogonal_center(4)
    plan = A.effective_hamiltonian_plan(4, *args, tol=1e-12)
    x = cpu_be.from_numpy(np.random.RandomState(0).randn(*A.mps[4].shape))
    e_before = float((x * plan.matvec(x)).sum())
    A.calculate_entanglement_spectrum()          # SVD re-gauge of every bond, centre returns to 4
    plan2 = A.effective_hamiltonian_plan(4, *args, tol=1e-12)
    # H_eff in the new gauge is a different matrix, but <psi|H|psi> of the state itself is gauge invariant
    psi = A.mps[4]
    e_state = float((psi * plan2.matvec(psi)).sum()) / float((psi * psi).sum())
    ref = g['p4_mps_4']
    assert abs(e_state - float(ref.reshape(-1) @ g['p4_heff'] @ ref.reshape(-1)) / float((ref ** 2).sum())) < 1e-12
    assert np.isfinite(e_before)


def test_term_table_rejects_unsorted_sites():
    from tnalg_b200.envs import TermTable
    ops = [np.eye(2)] * 7
    with pytest.raises(ValueError):
        TermTable(np.zeros((0, 2)), np.array([[3, 1, 3, 3]]), np.zeros(0), np.ones(1), ops, 1e-12)


@pytest.mark.parametrize('tag', ['chi16', 'chi24'])
def test_observables_of_reference_result_pickles(golden, cpu_be, tag):
    """host logic on the MPS stored in the reference's own data_dmrg/*.pr files (same check as the GPU test)"""
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    g = golden('pr_fixtures')
    L, d, chi = int(g[tag + '_para_l']), int(g[tag + '_para_d']), int(g[tag + '_para_chi'])
    ops_ = [np.real(o) if np.abs(np.imag(o)).max() == 0 else o for o in g[tag + '_para_op']]
    A = MpsOpenBoundaryClass(L, d, chi, operators=ops_)
    A.load_tensors([g['%s_mps_%d' % (tag, n)] for n in range(L)], int(g[tag + '_center']))
    assert np.abs(A.observe_magnetization(3).reshape(-1) - g[tag + '_mz'].reshape(-1)).max() < 1e-10
    eb = A.observe_bond_energy(g[tag + '_para_index2'], g[tag + '_para_coeff2'])
    assert np.abs(eb.reshape(-1) - g[tag + '_eb_full'].reshape(-1)).max() < 1e-10
    A.calculate_entanglement_spectrum()
    for n in range(L - 1):
        ref = g['%s_lm_%d' % (tag, n)]
        assert np.abs(A.lm[n] - ref).max() <= 1e-10 * ref.max() + 1e-13, n


def test_truncate_virtual_bonds_vs_oracle(cpu_be):
    """a12: SVD truncation of every bond to chi (library/MPSClass.py:186-247, 909-923)"""
    _check_truncation()


def _check_truncation():
    import os
    from oracle import dmrg_oracle as orc
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    # against the reference's library/ generation (golden) ...
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'truncation_lib.npz')))
    Lg = int(g['l'])
    G = MpsOpenBoundaryClass(Lg, int(g['d']), int(g['chi0']))
    G.load_tensors([g['mps_in_%d' % n] for n in range(Lg)], center=0)
    G.center = -1
    G.correct_orthogonal_center(0)
    G.truncate_virtual_bonds(int(g['chi1']), center=int(g['center']))
    assert np.array_equal(G.virtual_dim, g['virtual_dim'])
    for n in range(Lg - 1):
        assert np.abs(G.lm[n] - g['lm_%d' % n]).max() <= 1e-10 * g['lm_%d' % n].max(), n
    outg = [np.array(t.cpu().numpy()) for t in G.mps]
    assert abs(abs(orc.mps_overlap([g['mps_out_%d' % n] for n in range(Lg)], outg)) - 1) < 1e-10
    # ... and against the oracle on a fresh random state
    np.random.seed(3)
    L, d, chi0, chi1 = 8, 2, 12, 5
    A = MpsOpenBoundaryClass(L, d, chi0)
    host = [np.array(t.cpu().numpy()) for t in A.mps]
    ref, lms = orc.truncate_mps(host, chi1)
    A.correct_orthogonal_center(0)
    A.truncate_virtual_bonds(chi1, center=L - 1)
    got = [np.array(t.cpu().numpy()) for t in A.mps]
    assert max(A.virtual_dim) == chi1 and [t.shape for t in got] == [t.shape for t in ref]
    for n in range(L - 1):
        assert np.abs(A.lm[n] - lms[n]).max() <= 1e-10 * lms[n].max()           # kept spectrum
    assert abs(abs(orc.mps_overlap(ref, got)) - 1) < 1e-10                       # same state up to gauge/sign
    assert abs(A.norm_mps() - 1) < 1e-12
    # 'simple' way: plain index cut + re-orthogonalisation
    B = MpsOpenBoundaryClass(L, d, chi0)
    B.correct_orthogonal_center(0)
    B.truncate_virtual_bonds(chi1, center=3, way='simple')
    assert max(B.virtual_dim) <= chi1 and B.center == 3 and abs(B.norm_mps() - 1) < 1e-12


def _converged_state_for_rdm(golden):
    from tnalg_b200 import DMRG_anyH
    g = golden('rdm_xxz8')
    para = para_from_golden(g)
    np.random.seed(int(g['seed']))
    ob, A, info, para = DMRG_anyH.dmrg_finite_size(para)
    return g, para, ob, A


def check_rdm_and_dense_helpers(golden, be):
    """reduced_density_matrix_two_body vs the unmodified reference (MPSClass.py:841-855) for centres left of / inside /
    right of the pair, vs the partial trace of full_coefficients_mps; effective_hamiltonian_dmrg vs the oracle's dense
    H_eff; the check_* helpers.  Shared by the CPU (stand-in backend) and GPU tests."""
    from oracle import dmrg_oracle as orc
    g, para, ob, A = _converged_state_for_rdm(golden)
    assert abs(float(np.ravel(ob['e_per_site'])[0]) - float(g['e_per_site'][0])) <= 1e-10 * abs(float(g['e_per_site'][0]))
    A.mps = [be.from_numpy(np.asarray(t)) for t in A.mps]       # back on the device after clean_to_save
    L, d = para['l'], para['d']
    psi = None
    for c in (0, 4, 7):
        A.correct_orthogonal_center(c)
        assert A.check_orthogonality_by_tensors(is_print=False) == [] and A.check_virtual_bond_dimensions() == []
        assert abs(A.check_mps_norm1() - 1) < 1e-12
        if psi is None:
            psi = A.full_coefficients_mps().reshape([d] * L)
            assert abs(np.linalg.norm(psi) - 1) < 1e-12
        for p1, p2 in g['pairs']:
            rho = A.reduced_density_matrix_two_body(int(p1), int(p2))
            assert np.abs(rho - g['rdm_c%d_%d_%d' % (c, p1, p2)]).max() < 1e-8          # reference, observables tolerance
            keep = [int(p1), int(p2)]
            rest = [n for n in range(L) if n not in keep]
            m = psi.transpose(keep + rest).reshape(d * d, -1)
            assert np.abs(rho - m @ m.T).max() < 1e-12                                  # partial trace of the full state
    # dense H_eff through the plan == oracle's kron-built dense H_eff on the same tensors
    p = 3
    h = A.effective_hamiltonian_dmrg(p, para['index1'], para['index2'], para['coeff1'], para['coeff2'], tol=1e-12)
    host = [be.to_numpy(t) for t in A.mps]
    O = orc.OracleMps(L, d, para['chi'], [np.real(o) for o in para['op']], mps=host)
    O.center = A.center
    h0 = O.dense_effective_hamiltonian(p, para['index1'], para['index2'], para['coeff1'], para['coeff2'], 1e-12)
    assert np.abs(h - h0).max() < 1e-12 and np.abs(h - h.T).max() < 1e-12
    x = be.to_numpy(A.mps[p]).reshape(-1)
    assert abs(x @ h @ x - float(g['e_per_site'][0]) * L) < 1e-9


def test_rdm_full_state_dense_heff_and_checks(golden, cpu_be):
    check_rdm_and_dense_helpers(golden, cpu_be)


def check_tensor_helpers(golden, be):
    """a6 and friends against outputs of the reference's own functions (tests/golden/tensor_kats.npz) and against the
    integer examples printed in its docstrings (TensorBasicModule.py:476-485,434-441)"""
    from tnalg_b200 import TensorBasicModule as T
    g = golden('tensor_kats')
    M = [g['M0'], g['M1'], g['M2']]
    assert np.abs(T.absorb_matrices2tensor(g['T'], M) - g['absorb_all']).max() < 1e-12
    assert np.abs(T.absorb_matrices2tensor_full_fast(g['T'], M) - g['absorb_full_fast']).max() < 1e-12
    assert np.abs(T.absorb_matrices2tensor(g['T'], [M[2], M[0]], bonds=[2, 0]) - g['absorb_bonds_2_0']).max() < 1e-12
    ones = np.ones((2, 2, 2))
    m3 = [np.array([[1., 2], [2, 3]]), np.array([[2., 3], [3, 4]]), np.array([[3., 4], [4, 5]])]
    assert np.array_equal(np.rint(T.absorb_matrices2tensor_full_fast(ones, m3)), [[[105, 135], [147, 189]], [[175, 225], [245, 315]]])
    for name, fn, v2, v4 in (('l2r', T.bound_vec_with_phys_left2right, g['v2l'], g['v4l']),
                             ('r2l', T.bound_vec_with_phys_right2left, g['v2r'], g['v4r'])):
        assert np.abs(fn(g['T']) - g['phys_%s_empty' % name]).max() < 1e-12
        assert np.abs(fn(g['T'], v2) - g['phys_%s_v2' % name]).max() < 1e-12
        assert np.abs(fn(g['T'], v4) - g['phys_%s_v4' % name]).max() < 1e-12
    out = T.bound_vec_with_phys_left2right(be.from_numpy(g['T']), be.from_numpy(g['v2l']))   # device in -> device out
    assert hasattr(out, 'data_ptr') and np.abs(be.to_numpy(out) - g['phys_l2r_v2']).max() < 1e-12
    assert np.abs(T.transfer_matrix_mps(g['T']) - g['transfer_matrix']).max() < 1e-12
    tn, nrm = T.normalize_tensor(g['T'])
    assert np.abs(tn - g['normalized']).max() < 1e-15 and abs(nrm - float(g['norm'])) < 1e-13
    assert T.check_orthogonality(g['Q'], [2], tol=1e-12) == bool(g['check_ort_Q_2'])
    assert T.check_orthogonality(g['Q'], [0], tol=1e-12) == bool(g['check_ort_Q_0'])
    assert np.array_equal(T.ones_open_mps(3, 2, 3)[1], g['ones_mps_1'])


def test_tensor_helpers_vs_reference_functions(golden, cpu_be):
    check_tensor_helpers(golden, cpu_be)


def test_small_utilities_keep_reference_behaviour(capsys):
    """helpers next to the path with the reference's names: values checked against the reference's docstring examples
    (BasicFunctionsSJR.py:121-133,152-163,204-240,416-437; HamiltonianModule.py:60-72,137-148)"""
    from tnalg_b200 import BasicFunctionsSJR as B, HamiltonianModule as H, Parameters as Pm
    assert B.sort_list([1, 2, 'a', 'b'], [1, 3]) == [2, 'b']
    assert B.remove_element_from_list([1, 2, 3], 3) == [1, 2]
    assert B.arg_find_list([1, 2, 1, 3], 1, which='last') == [2] and B.arg_find_list([1, 2, 1, 3], 1, n=2) == [0, 2]
    assert B.print_sep('This is an example', '@', 20) == '@@@@@@@@@@ This is an example @@@@@@@@@@'
    assert B.print_dict({'name1': 1, 'name2': 'a'}, None, 'this is an example\n', '-') == 'this is an example\nname1-1\nname2-a\n'
    assert B.print_options(['left', 'right'], [1, 2], 'Where to go:') == 'Where to go:1: left    2: right'
    assert set(B.info_contact()) == {'name', 'email', 'affiliation'}
    pairs, n = H.interactions_full_connection_two_body(4)
    assert n == 6 and pairs.tolist() == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]
    assert H.from_spin2phys_dim('half') == 2 and H.from_spin2phys_dim('one') == 3
    h = H.hamiltonian_heisenberg('half', 1, 1, 1, 0, 0)
    assert np.allclose(np.linalg.eigvalsh(h), [-0.75, 0.25, 0.25, 0.25])          # singlet / triplet of two spins 1/2
    hz = H.hamiltonian_heisenberg('half', 0, 0, 0, 0, 0.5)
    assert np.allclose(np.diag(hz), [0.5, 0, 0, -0.5])
    assert 'The parameters are' in Pm.show_parameters({'l': 4})
    capsys.readouterr()


def _jacobi_rows_numpy(W, tol=5e-15, max_sweeps=60):
    """numpy emulation of tn_svd_jacobi's round-robin one-sided Jacobi on the ROWS of W (same pair schedule and rotation
    formula as the unblocked kernel); returns (sweeps, rotated W, accumulated rotations J with W_out = J W_in)"""
    W = W.copy()
    n = W.shape[0]
    J = np.eye(n)
    for sweep in range(max_sweeps):
        nrot = 0
        for rnd in range(n - 1):
            pq = []
            for i in range(n // 2):
                a, b = (rnd % (n - 1), n - 1) if i == 0 else ((rnd + i) % (n - 1), (rnd + n - 1 - i) % (n - 1))
                pq.append((min(a, b), max(a, b)))
            P, Q = np.array([p for p, _ in pq]), np.array([q for _, q in pq])
            X, Y = W[P], W[Q]
            a, b, g = (X * X).sum(1), (Y * Y).sum(1), (X * Y).sum(1)
            rot = (a > 0) & (b > 0) & (g * g > tol * tol * a * b)
            if not rot.any():
                continue
            nrot += int(rot.sum())
            gg = np.where(rot, g, 1.0)
            zeta = (b - a) / (2 * gg)
            t = np.where(zeta >= 0, 1.0, -1.0) / (np.abs(zeta) + np.sqrt(1 + zeta ** 2))
            c = 1 / np.sqrt(1 + t * t)
            s = c * t
            c, s = np.where(rot, c, 1.0)[:, None], np.where(rot, s, 0.0)[:, None]
            W[P], W[Q] = c * X - s * Y, s * X + c * Y
            JP, JQ = J[P], J[Q]
            J[P], J[Q] = c * JP - s * JQ, s * JP + c * JQ
        if nrot == 0:
            return sweep + 1, W, J
    return max_sweeps, W, J


def test_preconditioned_svd_algebra_and_sweep_counts(golden):
    """ops.preconditioned_svd (the code the CUDA backend runs, here with numpy callables): exact truncated SVD on random and
    on real two-site wavefunctions, and the norm-sorted QR steps cut the Jacobi sweeps on the latter"""
    import torch
    from tnalg_b200.ops import preconditioned_svd
    sweeps = []

    def qr(X):
        q, r = np.linalg.qr(X.numpy())
        return torch.from_numpy(np.ascontiguousarray(q)), torch.from_numpy(np.ascontiguousarray(r))

    def jacobi(X, k):                      # X (n, n): rows of X^T are rotated, as the kernel does for tall inputs
        n_sw, W, J = _jacobi_rows_numpy(X.numpy().T.copy())
        sweeps.append(n_sw)
        sig = np.linalg.norm(W, axis=1)
        order = np.argsort(-sig)[:k]
        U = (W[order] / sig[order, None]).T                      # X = U S Vt with W = J X^T = S U^T
        return (torch.from_numpy(np.ascontiguousarray(U)), torch.from_numpy(sig[order].copy()),
                torch.from_numpy(np.ascontiguousarray(J[order])))

    def mm_nn(X, Y):
        return X @ Y

    def mm_nt(X, Y):
        return X @ Y.t()

    rng = np.random.RandomState(3)
    cases = [rng.randn(80, 64), (np.linalg.qr(rng.randn(70, 70))[0] * np.logspace(0, -10, 70)) @ np.linalg.qr(rng.randn(70, 70))[0].T]
    cases += list(golden('theta_two_site')['theta'])
    for n_case, A in enumerate(cases):
        m, n = A.shape
        for k in (n, n // 2):
            U, S, Vt = [x.numpy() for x in preconditioned_svd(torch.from_numpy(A.copy()), k, qr, jacobi, mm_nn, mm_nt)]
            u0, s0, v0 = np.linalg.svd(A, full_matrices=False)
            assert U.shape == (m, k) and Vt.shape == (k, n)
            assert np.abs(S - s0[:k]).max() < 1e-13 * s0[0]
            assert np.abs(U.T @ U - np.eye(k)).max() < 1e-12 and np.abs(Vt @ Vt.T - np.eye(k)).max() < 1e-12
            assert np.abs((U * S) @ Vt - (u0[:, :k] * s0[:k]) @ v0[:k]).max() < 1e-12 * s0[0]
    # sweep counts on the real wavefunctions: sorted preconditioning vs the plain double QR
    for A in golden('theta_two_site')['theta']:
        sweeps.clear()
        preconditioned_svd(torch.from_numpy(A.copy()), A.shape[1], qr, jacobi, mm_nn, mm_nt)
        sorted_sweeps = sweeps[-1]
        r2 = np.linalg.qr(np.linalg.qr(A)[1].T)[1]
        plain_sweeps = _jacobi_rows_numpy(r2.copy())[0]           # Jacobi on the columns of R2^T = rows of R2
        assert sorted_sweeps <= 9 and sorted_sweeps + 3 <= plain_sweeps, (sorted_sweeps, plain_sweeps)


def test_cuda_backend_svd_wrapper_with_stand_in_primitives():
    """the body of ops.CudaBackend.svd (shape handling, transposed input, truncation, the GEMM wrappers and their leading
    dimensions) executed on the CPU with stand-ins for the four primitives it calls (`empty`, `_gemm`, `qr`, `_jacobi`)"""
    import torch
    from tnalg_b200 import ops

    class Fake:
        last_svd_sweeps = 0

        def empty(self, *shape):
            return torch.full(shape, float('nan'), dtype=torch.float64)

        def _gemm(self, mode, M, N, K, A, B, lda, ldb, Cout, deterministic=1):
            assert A.is_contiguous() and B.is_contiguous() and Cout.shape == (M, N)
            if mode == 0:      # C = A (M x K, pitch lda) . B (K x N, pitch ldb)
                assert A.shape == (M, K) and B.shape == (K, N) and lda == K and ldb == N
                Cout.copy_(A @ B)
            elif mode == 1:    # C = A (M x K) . B (N x K)^T
                assert A.shape == (M, K) and B.shape == (N, K) and lda == K and ldb == K
                Cout.copy_(A @ B.t())
            else:
                raise AssertionError(mode)
            return Cout

        def qr(self, A):
            q, r = torch.linalg.qr(A, mode='reduced')
            return q.t().contiguous().t(), r.t().contiguous().t()      # column-major like cuSOLVER's outputs

        def _jacobi(self, A, k_keep=None):
            u, s, vt = torch.linalg.svd(A, full_matrices=False)
            k = s.numel() if k_keep is None else k_keep
            return u[:, :k].contiguous(), s[:k].contiguous(), vt[:k].contiguous()

        svd = ops.CudaBackend.svd

    fake = Fake()
    rng = np.random.RandomState(5)
    for (m, n) in [(96, 80), (80, 96), (128, 128), (40, 30), (30, 40), (5, 1), (1, 5)]:
        A = rng.randn(m, n) * np.logspace(0, -6, n)[None, :]
        for k in (None, max(1, min(m, n) // 2)):
            for pre in (True, False):
                U, S, Vt = [x.numpy() for x in fake.svd(torch.from_numpy(A.copy()), k_keep=k, precondition=pre)]
                kk = min(m, n) if k is None else k
                u0, s0, v0 = np.linalg.svd(A, full_matrices=False)
                assert U.shape == (m, kk) and S.shape == (kk,) and Vt.shape == (kk, n)
                assert not np.isnan(U).any() and not np.isnan(Vt).any()
                assert np.abs(S - s0[:kk]).max() < 1e-12 * s0[0]
                assert np.abs((U * S) @ Vt - (u0[:, :kk] * s0[:kk]) @ v0[:kk]).max() < 1e-11 * s0[0]


@pytest.mark.parametrize('case', ['e2e_full6', 'e2e_jigsaw7'])
def test_energy_parity_on_degenerate_lattices(golden, cpu_be, case):
    """all-to-all Heisenberg ('full') and the spin-1 'jigsaw' chain have degenerate ground multiplets: the converged energy
    is a parity target (1e-10), the observables inside the multiplet are not"""
    from tnalg_b200.DMRG_anyH import dmrg_finite_size
    g = golden(case)
    para = para_from_golden(g)
    np.random.seed(int(g['seed']))
    ob, A, info, para = dmrg_finite_size(para)
    assert abs(ob['e_per_site'][0] - g['e_per_site'][0]) <= 1e-10 * abs(g['e_per_site'][0])
    assert np.array_equal(A.virtual_dim, g['virtual_dim']) and info['not_converged'] == 0


def test_lattice_generators_match_reference_tables(golden):
    """Parameters.generate_parameters_dmrg / make_consistent_parameter_dmrg for every lattice the reference generates
    (Parameters.py:64-165,188-277): same index and coefficient tables as the unmodified reference"""
    from tnalg_b200 import Parameters as Pm
    cases = {'e2e_longrange8': ('longRange', dict(l=8, chi=16, jxy=0, jz=1, hx=0.5, hz=0, alpha=1.0)),
             'e2e_square3x2': ('square', dict(square_width=3, square_height=2, chi=8)),
             'e2e_full6': ('full', dict(l=6, chi=8)),
             'e2e_jigsaw7': ('jigsaw', dict(l=7, chi=27))}
    for name, (lattice, kw) in cases.items():
        g = golden(name)
        para = Pm.generate_parameters_dmrg(lattice)
        para.update(kw)
        if lattice == 'square':
            para['op'] = para['op'][:6]   # like the reference, the generator appends the field operator on every call
        para = Pm.make_consistent_parameter_dmrg(para)
        assert para['l'] == int(g['l']) and para['d'] == int(g['d']), name
        for k in ('index1', 'index2', 'positions_h2'):
            assert np.array_equal(np.asarray(para[k]), g[k]), (name, k)
        for k in ('coeff1', 'coeff2'):
            assert np.allclose(np.asarray(para[k], dtype=float).reshape(-1), g[k].reshape(-1), rtol=0, atol=1e-15), (name, k)
        assert len(para['op']) == g['op'].shape[0] and all(np.allclose(a, b) for a, b in zip(para['op'], g['op'])), name


def test_environment_cache_is_keyed_on_content(cpu_be):
    """an in-place edit of para['coeff2'] must not leave stale environments (the reference rebuilds them
    from coeff* on every update_tensor_eigs call, MPSClass.py:633-682) -- the reproduction of the advisor, on the stand-in"""
    from oracle import dmrg_oracle as orc
    from tnalg_b200 import Parameters as Pm
    from tnalg_b200.DMRG_anyH import observe, sweep_once
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=6, chi=8, eigs_tol=1e-12)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(0)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(0)
    for _ in range(3):
        sweep_once(A, para)
    env_before = A._env
    sweep_once(A, para)
    assert A._env is env_before                     # unchanged couplings: the cache is reused
    para['coeff2'][2::3] *= 3.0
    for _ in range(6):
        sweep_once(A, para)
    assert A._env is not env_before
    e = observe(A, para, {})['e_per_site'][0] * para['l']
    e_ed = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert abs(e - e_ed) < 1e-9 * abs(e_ed)
    # an operator edited in place (the field term op[6] after hx changes) is seen as well
    env_before = A._env
    para['op'][6][:] = -0.4 * np.real(para['op'][1])
    sweep_once(A, para)
    assert A._env is not env_before


def test_eig_way_0_and_complex_operator_errors(cpu_be):
    from oracle import dmrg_oracle as orc
    from tnalg_b200 import Parameters as Pm
    from tnalg_b200.DMRG_anyH import dmrg_finite_size
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=6, chi=8, eigs_tol=1e-12, break_tol=1e-13, eigWay=0, dt_ob=1)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(0)
    ob, A, info, _ = dmrg_finite_size(para)
    e_ed = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert abs(ob['e_per_site'][0] * para['l'] - e_ed) < 1e-9 * abs(e_ed)
    np.random.seed(0)
    B = MpsOpenBoundaryClass(4, 2, 4)
    B.correct_orthogonal_center(1)
    with pytest.raises(ValueError, match='complex'):
        B.observe_magnetization(2)
    B.eig_way = 0
    B.mps[1] = np.random.randn(4, 2, 4)
    with pytest.raises(ValueError, match='eig_way=0'):
        B._dense_solve(None, (64, 2, 64), None, 1e-4)


def test_eigs_fh_host_logic(cpu_be):
    """a8' call shape on the stand-in: callable lin_map, n > 1 by deflation, which = sa / la / lm"""
    from tnalg_b200.Eigs_Module_sjr import eigs_fh
    rng = np.random.RandomState(0)
    d = 24
    g = rng.randn(d, d)
    H = (g + g.T) / 2
    w, V = np.linalg.eigh(H)
    lm, v, info = eigs_fh(lambda x: H @ x.numpy(), d, n=3, which='sa', tol=1e-12)
    assert lm.shape == (3,) and v.shape == (d, 3) and np.abs(lm - w[:3]).max() < 1e-8
    lm, v, info = eigs_fh(lambda x: H @ x.numpy(), d, n=2, which='la')
    assert np.abs(lm - w[::-1][:2]).max() < 1e-8
    lm, v, info = eigs_fh(lambda x: H @ x.numpy(), d, n=1, which='lm')
    assert abs(abs(lm[0]) - np.abs(w).max()) < 1e-7
    assert set(info) >= {'it_time', 'error'}


def test_sharded_observables_partition_covers_every_term_once(cpu_be):
    """envs.expect_products with a communicator: every rank contracts a block of leading sites and the all-reduce of the value
    vectors completes the result.  The ranks are simulated one after the other (the all-reduce of the stand-in communicator does
    nothing), so the partial vectors must have disjoint supports and add up to the unsharded values, for every world size."""
    from tnalg_b200 import envs
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    from tnalg_b200 import Parameters as Pm
    para = Pm.generate_parameters_dmrg('square')
    para.update(square_width=3, square_height=2, chi=8, op=para['op'][:6])
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(5)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(2)
    index2 = np.asarray(para['index2'], dtype=int)
    terms = [tuple(sorted(((int(r[0]), int(r[2])), (int(r[1]), int(r[3]))))) for r in index2]
    terms += [((i, 1),) for i in range(para['l'])] + [((i, 3),) for i in range(para['l'])]
    ops_real = [np.real(np.asarray(o)).astype(float) for o in para['op']]
    want = envs.expect_products(cpu_be, A.mps, A.center, ops_real, terms)

    class OneRank:
        def __init__(self, rank, world):
            self.rank, self.world = rank, world

        def allreduce(self, t):
            return t

    for world in (2, 3, 8):
        parts = [envs.expect_products(cpu_be, A.mps, A.center, ops_real, terms, comm=OneRank(r, world), min_terms=1)
                 for r in range(world)]
        support = np.sum([np.asarray(p) != 0 for p in parts], axis=0)
        assert support.max() <= 1
        assert np.abs(np.sum(parts, axis=0) - want).max() < 1e-13


def test_sharded_bond_move_places_every_block_once(cpu_be):
    """EnvCache._env_update with a communicator: the outgoing blocks are dealt round-robin (heaviest first), every rank writes its
    own into its rows of one buffer and an in-place all-gather completes it.  Ranks are simulated one after the other with an
    all-gather that does nothing: every block must be produced by exactly one rank, equal to the unsharded result, and land in
    the row all ranks agree on."""
    import torch
    from tnalg_b200 import envs
    rng = np.random.RandomState(11)
    a, d, b = 6, 2, 5
    T = torch.from_numpy(rng.randn(a, d, b))
    sz = np.array([[0.5, 0.0], [0.0, -0.5]])
    sp = np.array([[0.0, 1.0], [0.0, 0.0]])
    E = [torch.from_numpy(rng.randn(a, a)) for _ in range(4)]
    outputs = [[(E[0], None), (E[1], sz), (None, sz)], [(None, sp)], [(E[2], None)], [(E[3], sp)], [(None, sz)], [(E[1], None)], [(E[0], sp)]]
    want = cpu_be.env_update(0, T, outputs)

    class OneRank:
        def __init__(self, rank, world):
            self.rank, self.world = rank, world

        def allgather_inplace(self, buf):
            self.buf = buf
            return buf

    terms = envs.TermTable(np.zeros((0, 2), dtype=int), np.array([[0, 1, 1, 1]]), np.zeros((0, 1)), np.ones((1, 1)), [np.eye(2), sz, sp, sz], 1e-12)
    for world in (2, 3, 8):
        got = [None] * len(outputs)
        rows = None
        for r in range(world):
            cache = envs.EnvCache(cpu_be, terms, 2)
            cache.shard_min_dim = 1
            cache.comm = OneRank(r, world)
            res = cache._env_update(0, T, outputs)
            per = (len(outputs) + world - 1) // world
            assert cache.comm.buf.shape == (world * per, b, b)
            # the row of every block inside the buffer is the same on all ranks
            place = [int((blk.data_ptr() - cache.comm.buf.data_ptr()) // (8 * b * b)) for blk in res]
            rows = place if rows is None else rows
            assert place == rows and sorted(place) == sorted(set(place))
            for j, blk in enumerate(res):
                if r * per <= place[j] < (r + 1) * per:      # this rank's own rows
                    assert got[j] is None
                    got[j] = blk.clone()
        for j in range(len(outputs)):
            assert got[j] is not None and np.abs(got[j].numpy() - want[j].numpy()).max() < 1e-14
