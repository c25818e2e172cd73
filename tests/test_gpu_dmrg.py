"""End-to-end parity of the CUDA path (through the drop-in API) with the reference goldens -- run on the B200 box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize('p', [0, 2, 4, 5, 8])
def test_matvec_on_reference_snapshot(golden, p):
    """identical MPS -> the CUDA environments + matvec reproduce the reference handle output (a1..a5)"""
    from tnalg_b200 import ops
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    be = ops.backend()
    g = golden('percall_j1j2')
    L, d, chi = int(g['l']), int(g['d']), int(g['chi'])
    A = MpsOpenBoundaryClass(L, d, chi, operators=[np.real(o) for o in g['op']], is_save_op=True, eig_way=1)
    for n in range(L):
        A.mps[n] = g['p%d_mps_%d' % (p, n)]
    A.center = p
    plan = A.effective_hamiltonian_plan(p, g['index1'], g['index2'], g['coeff1'], g['coeff2'], tol=1e-12)
    x = be.from_numpy(g['p%d_x' % p].reshape(tuple(g['p%d_shape' % p])))
    y = be.to_numpy(plan.matvec(x, 1.0, -float(g['tau']))).reshape(-1)
    assert np.abs(y - g['p%d_y' % p]).max() < 1e-13 * max(1.0, np.abs(g['p%d_y' % p]).max())
    hx = be.to_numpy(plan.matvec(x, 0.0, 1.0)).reshape(-1)
    ref = g['p%d_heff' % p] @ g['p%d_x' % p]
    assert np.abs(hx - ref).max() < 1e-13 * np.abs(ref).max()


def test_observables_on_reference_snapshot(golden):
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    g = golden('percall_j1j2')
    L, d, chi = int(g['l']), int(g['d']), int(g['chi'])
    A = MpsOpenBoundaryClass(L, d, chi, operators=[np.real(o) for o in g['op']])
    for n in range(L):
        A.mps[n] = g['ob_mps_%d' % n]
    A.center = 4
    assert np.abs(A.observe_magnetization(1) - g['ob_mx']).max() < 1e-12
    assert np.abs(A.observe_magnetization(3) - g['ob_mz']).max() < 1e-12
    assert np.abs(A.observe_bond_energy(g['index2'], g['coeff2']) - g['ob_eb_full']).max() < 1e-12
    assert np.abs(A.observe_correlators_from_middle(3, 3) - g['ob_corr_z']).max() < 1e-12
    assert np.abs(A.observe_correlators_from_middle(1, 1) - g['ob_corr_x']).max() < 1e-12
    for c in (0, L - 1):
        A.correct_orthogonal_center(c)
        assert np.abs(A.observe_magnetization(3) - g['ob_mz']).max() < 1e-11


@pytest.mark.parametrize('case', ['e2e_chain12', 'e2e_xxz10', 'e2e_j1j2_4x2', 'e2e_spin1_chain8', 'e2e_longrange8', 'e2e_square3x2',
                                  'e2e_periodic8'])
def test_end_to_end_vs_reference(golden, case):
    """converged, tight-tolerance runs: sweep energies and truncated spectrum rel 1e-10, observables abs 1e-8"""
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
    assert all(isinstance(t, np.ndarray) for t in A.mps)


@pytest.mark.parametrize('case', ['e2e_full6', 'e2e_jigsaw7'])
def test_energy_parity_on_degenerate_lattices(golden, case):
    """all-to-all and spin-1 jigsaw lattices have degenerate ground multiplets: only the energy is a parity quantity"""
    from tnalg_b200.DMRG_anyH import dmrg_finite_size
    g = golden(case)
    para = para_from_golden(g)
    np.random.seed(int(g['seed']))
    ob, A, info, para = dmrg_finite_size(para)
    assert abs(ob['e_per_site'][0] - g['e_per_site'][0]) <= 1e-10 * abs(g['e_per_site'][0])
    assert info['not_converged'] == 0


def test_size_independent_properties_chi64():
    """larger than the oracle can check quickly: properties the domain offers.
    (1) the matvec is symmetric <x|H y> = <y|H x>; (2) the Lanczos energy is variational and non-increasing along
    the sweep; (3) the state stays normalised; (4) observables computed with the centre at both ends agree."""
    from tnalg_b200 import Parameters as Pm, ops
    from tnalg_b200.DMRG_anyH import sweep_once
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    be = ops.backend()
    para = Pm.generate_parameters_dmrg('square')
    para.update(square_width=4, square_height=4, chi=64, op=para['op'][:6])
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(4)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(0)
    sweep_once(A, para)
    A.correct_orthogonal_center(7)
    plan = A.effective_hamiltonian_plan(7, para['index1'], para['index2'], para['coeff1'], para['coeff2'], tol=1e-5)
    rng = np.random.RandomState(0)
    x, y = be.from_numpy(rng.randn(*A.mps[7].shape)), be.from_numpy(rng.randn(*A.mps[7].shape))
    xhy = float((x * plan.matvec(y)).sum())
    yhx = float((y * plan.matvec(x)).sum())
    assert abs(xhy - yhx) < 1e-11 * max(abs(xhy), 1.0)
    energies = []
    for _ in range(3):
        sweep_once(A, para)
        eb = A.observe_bond_energy(para['index2'], para['coeff2'])
        energies.append(float(eb.sum()))
        assert abs(A.norm_mps() - 1) < 1e-12
    assert energies[1] <= energies[0] + 1e-9 and energies[2] <= energies[1] + 1e-9
    mz0 = A.observe_magnetization(3)
    A.correct_orthogonal_center(para['l'] - 1)
    assert np.abs(A.observe_magnetization(3) - mz0).max() < 1e-10
    assert abs(float(A.observe_bond_energy(para['index2'], para['coeff2']).sum()) - energies[-1]) < 1e-9


@pytest.mark.parametrize('tag', ['chi16', 'chi24'])
def test_observables_of_reference_result_pickles(golden, tag):
    """the MPS stored in the reference's own data_dmrg/*.pr files: the CUDA observable path reproduces the stored
    magnetisation, bond energies and (via Jacobi SVD) the stored entanglement spectrum"""
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    g = golden('pr_fixtures')
    L, d, chi = int(g[tag + '_para_l']), int(g[tag + '_para_d']), int(g[tag + '_para_chi'])
    ops_ = [np.real(o) if np.abs(np.imag(o)).max() == 0 else o for o in g[tag + '_para_op']]
    A = MpsOpenBoundaryClass(L, d, chi, operators=ops_)
    A.load_tensors([g['%s_mps_%d' % (tag, n)] for n in range(L)], int(g[tag + '_center']))
    assert np.abs(A.observe_magnetization(3).reshape(-1) - g[tag + '_mz'].reshape(-1)).max() < 1e-10
    assert np.abs(A.observe_magnetization(1).reshape(-1) - g[tag + '_mx'].reshape(-1)).max() < 1e-10
    eb = A.observe_bond_energy(g[tag + '_para_index2'], g[tag + '_para_coeff2'])
    assert np.abs(eb.reshape(-1) - g[tag + '_eb_full'].reshape(-1)).max() < 1e-10
    assert abs(eb.sum() / L - float(g[tag + '_e_per_site'].reshape(-1)[0])) < 1e-10
    A.calculate_entanglement_spectrum()
    A.calculate_entanglement_entropy()
    for n in range(L - 1):
        ref = g['%s_lm_%d' % (tag, n)]
        assert np.abs(A.lm[n] - ref).max() <= 1e-10 * ref.max() + 1e-13, n
    assert np.abs(A.ent.reshape(-1) - g[tag + '_ent'].reshape(-1)).max() < 1e-9


def test_truncate_virtual_bonds_vs_oracle():
    """a12 on the device: Jacobi SVD with k_keep as the truncating gauge move"""
    from tests.test_host_logic_cpu import _check_truncation
    _check_truncation()


def _full_size_checks(chi, n_ranks):
    """size-independent properties on the bench workload (6x6 J1-J2): see test_full_size_properties_chi1024"""
    import bench
    from tnalg_b200 import ops
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    be = ops.backend()
    para = bench.build_para(bench.WORKLOADS['j1j2_6x6_chi1024'], chi)
    L, d = para['l'], para['d']
    args = (para['index1'], para['index2'], para['coeff1'], para['coeff2'])
    np.random.seed(8)
    A = MpsOpenBoundaryClass(L, d, para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    p = L // 2 - 1 if chi < 64 else 11                                   # widest site of the bench (a = b = chi)
    A.correct_orthogonal_center(p)
    A.mps[p] = A.mps[p] / A.norm_mps()                                    # random tensors: normalise the centre
    assert abs(A.norm_mps() - 1) < 1e-12
    plan = A.effective_hamiltonian_plan(p, *args, tol=para['eigs_tol'])
    psi = A.mps[p]
    rng = np.random.RandomState(1)
    x, y = be.from_numpy(rng.randn(*psi.shape)), be.from_numpy(rng.randn(*psi.shape))
    hx, hy = plan.matvec(x).clone(), plan.matvec(y).clone()
    # (1) symmetry and (2) linearity of the matvec
    xhy, yhx = float((x * hy).sum()), float((y * hx).sum())
    assert abs(xhy - yhx) < 1e-11 * float(hx.norm()) * float(y.norm())
    hz = plan.matvec(0.3 * x - 1.7 * y)
    assert float((hz - (0.3 * hx - 1.7 * hy)).norm()) < 1e-12 * float(hx.norm() + hy.norm())
    # (3) term-sharded plans add up to the full operator (the multi-GPU decomposition, without a collective)
    parts = None
    for r in range(n_ranks):
        pr = A._environments(*args, para['eigs_tol']).plan(p, A.mps, rank=r, world=n_ranks)
        out = pr.matvec(x).clone()
        parts = out if parts is None else parts + out
        pr.destroy()
    assert float((parts - hx).norm()) < 1e-12 * float(hx.norm())
    # (4) checksum of checksums: <psi|H_eff|psi> through the matvec kernels == sum of all bond energies through the
    #     observable path (independent code: expect_products / tn_env_update chains / tn_trace)
    e_plan = float((psi * plan.matvec(psi)).sum())
    e_obs = float(np.sum(A.observe_bond_energy(para['index2'], para['coeff2'])))
    assert abs(e_plan - e_obs) < 1e-10 * max(1.0, abs(e_obs)), (e_plan, e_obs)
    # (5) one local solve is variational and leaves a normalised, gauge-consistent state
    A.update_tensor_eigs(p, *args, para['tau'], para['is_real'], tol=para['eigs_tol'])
    assert abs(A.norm_mps() - 1) < 1e-12
    e_after = float(np.sum(A.observe_bond_energy(para['index2'], para['coeff2'])))
    assert e_after <= e_obs + 1e-9 * abs(e_obs)
    assert A.last_eig['converged'] and abs((1.0 - A.last_eig['lambda']) / para['tau'] - e_after) < 1e-6 * abs(e_after)
    plan.destroy()
    return A, para, p


def _oracle_handle_from_blocks(A, para, p, x):
    """the reference handle (oracle.apply_handle, MPSClass.py:755-776) evaluated on the host with the UNMERGED opt_env groups
    ('1_0_1' as the literal list of crossing terms) copied from the device blocks"""
    from oracle import dmrg_oracle as orc
    from tnalg_b200 import ops
    be = ops.backend()
    env = A._environments(para['index1'], para['index2'], para['coeff1'], para['coeff2'], para['eigs_tol'])
    env.ensure(p, A.mps)
    a, d, b = A.mps[p].shape
    env.merge_crossing = False
    try:
        g = env.groups(p, d)
    finally:
        env.merge_crossing = True
    h = lambda t: be.to_numpy(t)  # noqa: E731
    site_ops = [np.eye(d)] + [np.asarray(o) for o in g['ls_ops']] + [np.asarray(o) for o in g['rs_ops']]   # index 0 is reserved ('1_0_0')
    oenv = {}
    if g['HL'] is not None:
        oenv['1_0_0'] = h(g['HL'])
    if g['HR'] is not None:
        oenv['0_0_1'] = h(g['HR'])
    if g['M'] is not None:
        oenv['0_99_0'] = np.asarray(g['M'])
    for k, m in enumerate(g['LS']):
        oenv['1_%d_0' % (k + 1)] = h(m)
    for k, m in enumerate(g['RS']):
        oenv['0_%d_1' % (len(g['LS']) + k + 1)] = h(m)
    if g['XL']:
        oenv['1_0_1'] = [[c, h(l), h(r)] for c, l, r in zip(g['x_coeff'], g['XL'], g['XR'])]
    O = orc.OracleMps(2, d, 2, site_ops, mps=[np.zeros((1, d, 2)), np.zeros((2, d, 1))])
    return O.apply_handle(x.reshape(-1), oenv, (a, d, b), para['tau']), len(g['XL'])


def test_full_size_properties_chi1024():
    """BASELINE.json's full size (6x6 J1-J2, chi = 1024, a = b = 1024, K_L = K_R = 4, n_x = 42 at the probed site 11):
    (i) size-independent properties -- symmetry, linearity, shard additivity, the energy of the state computed by the matvec
    path against the same energy from the observable path, a variational local update;
    (ii) DIRECT parity at the metric shape: the merged 15-link CUDA plan against the reference handle evaluated by the oracle
    on the unmerged 42 crossing terms (1e-13 relative), and one left-to-right and one right-to-left environment update
    against oracle.transfer_l2r / transfer_r2l at chi = 1024"""
    from oracle import dmrg_oracle as orc
    from tnalg_b200 import ops
    be = ops.backend()
    A, para, p = _full_size_checks(1024, 4)
    shape = tuple(A.mps[p].shape)
    assert shape == (1024, 2, 1024)
    rng = np.random.RandomState(5)
    x = rng.randn(*shape)
    plan = A.effective_hamiltonian_plan(p, para['index1'], para['index2'], para['coeff1'], para['coeff2'], tol=para['eigs_tol'])
    got = be.to_numpy(plan.matvec(be.from_numpy(x), 1.0, -para['tau'])).reshape(-1)
    hx = be.to_numpy(plan.matvec(be.from_numpy(x), 0.0, 1.0)).reshape(-1)
    plan.destroy()
    ref, n_cross = _oracle_handle_from_blocks(A, para, p, x)
    assert n_cross == 42
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
    href = (x.reshape(-1) - ref) / para['tau']          # H x from the reference handle (loses ~4 digits to the shift)
    assert np.abs(hx - href).max() <= 1e-9 * np.abs(href).max()
    # environment updates at chi = 1024 (a5)
    T = A.mps[p]
    Tn = be.to_numpy(T)
    E = rng.randn(1024, 1024)
    op = np.array([[0.3, -1.2], [0.7, 0.4]])
    for direction, fn in ((0, orc.transfer_l2r), (1, orc.transfer_r2l)):
        out = be.to_numpy(be.env_update(direction, T, [[(be.from_numpy(E), op)]])[0])
        want = fn(Tn, op, E)
        assert np.abs(out - want).max() <= 1e-13 * np.abs(want).max(), direction


def test_rdm_full_state_dense_heff_and_checks(golden):
    """a10 reduced_density_matrix_two_body (vs the unmodified reference), full_coefficients_mps, a9 dense H_eff through the
    matvec plan (vs the oracle) and the check_* helpers on the CUDA backend"""
    from tests.test_host_logic_cpu import check_rdm_and_dense_helpers
    from tnalg_b200 import ops
    check_rdm_and_dense_helpers(golden, ops.backend())
