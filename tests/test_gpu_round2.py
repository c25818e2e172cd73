"""Round-2 entry points on the CUDA path (through the C ABI), each against the numpy oracle / LAPACK on identical seeded inputs:
Householder QR (a7), Jacobi eigh + dense eig_way=0 (a9/f4), ED on the full space (f4), eigs_fh (a8'), the multi-matrix wrappers
of TensorBasicModule (a6), observables at chi = 512 (a10), row-sliced plans (e), deterministic mode, spin-1 two-site (f1)."""
import numpy as np
import pytest

from oracle import dmrg_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def be():
    from tnalg_b200 import ops
    return ops.backend()


def graded(rng, m, n, cond):
    a = rng.randn(m, n)
    if cond > 1:
        u, s, vt = np.linalg.svd(a, full_matrices=False)
        a = (u * np.logspace(0, -np.log10(cond), s.size)) @ vt
    return a


QR_SHAPES = [(1, 1, 1), (2, 2, 1), (2, 16, 1), (4, 2, 1), (33, 31, 1), (64, 32, 1), (96, 96, 1e8), (200, 100, 1e12), (257, 130, 1e15),
             (50, 120, 1), (512, 256, 1e10), (700, 64, 1), (1024, 512, 1e12), (2048, 1024, 1e6), (300, 300, 1e13)]


@pytest.mark.parametrize('m,n,cond', QR_SHAPES)
def test_householder_qr_vs_lapack(be, m, n, cond):
    """a7: A = Q R with Q^T Q = 1 to 1e-13 for kappa up to 1e15, R equal to np.linalg.qr's including row signs (same reflector
    convention as LAPACK), Q R = A to rounding"""
    rng = np.random.RandomState(m * 7 + n)
    a = graded(rng, m, n, cond)
    Q, R = be.qr(be.from_numpy(a))
    q, r = be.to_numpy(Q), be.to_numpy(R)
    k = min(m, n)
    assert q.shape == (m, k) and r.shape == (k, n)
    assert np.abs(q.T @ q - np.eye(k)).max() < 1e-13
    assert np.abs(q @ r - a).max() <= 1e-14 * max(m, n) ** 0.5 * np.abs(a).max()
    assert np.abs(np.tril(r, -1)).max() == 0.0
    q0, r0 = np.linalg.qr(a)
    scale = np.abs(r0).max()
    # rows of R whose diagonal is well above round-off are unique up to sign; ours must match LAPACK's sign convention
    good = np.abs(np.diag(r0)) > 1e-9 * scale
    assert np.abs(r[:k][good] - r0[good]).max() <= 1e-11 * scale * max(1.0, min(cond, 1e4) ** 0.0)
    assert np.abs(np.abs(r) - np.abs(r0)).max() <= 2e-10 * scale or not good.all()


def test_qr_rank_deficient_and_zero_columns(be):
    rng = np.random.RandomState(3)
    a = rng.randn(96, 64)
    a[:, 32:] = a[:, :32]                  # exactly dependent columns
    a[:, 5] = 0.0                          # a zero column
    Q, R = be.qr(be.from_numpy(a))
    q, r = be.to_numpy(Q), be.to_numpy(R)
    assert np.abs(q.T @ q - np.eye(64)).max() < 1e-13       # Householder: an isometry whatever the rank
    assert np.abs(q @ r - a).max() < 1e-13 * np.abs(a).max()
    ones = np.ones((40, 2, 40))                              # ini_way='1' tensors: rank one
    Qt, Rt = be.qr_tensor(be.from_numpy(ones), True)
    qq = be.to_numpy(Qt).reshape(80, 40)
    assert np.abs(qq.T @ qq - np.eye(40)).max() < 1e-13 and np.abs(qq @ be.to_numpy(Rt) - ones.reshape(80, 40)).max() < 1e-12


@pytest.mark.parametrize('shape', [(1, 2, 16), (16, 2, 1), (8, 2, 16), (16, 2, 16), (64, 3, 64), (256, 2, 256), (37, 2, 53), (100, 4, 30)])
def test_qr_gauge_moves_l2r_r2l_vs_oracle(be, shape):
    """tn_qr_l2r / tn_qr_r2l against left2right/right2left_decompose_tensor of the oracle (TensorBasicModule.py:314-384)"""
    rng = np.random.RandomState(sum(shape))
    t = rng.randn(*shape)
    a, d, b = shape
    Q, R = be.qr_tensor(be.from_numpy(t), True)
    q, r = be.to_numpy(Q), be.to_numpy(R)
    k = min(a * d, b)
    assert q.shape == (a, d, k) and r.shape == (k, b)
    assert np.abs(np.einsum('asx,asy->xy', q, q) - np.eye(k)).max() < 1e-13
    assert np.abs(np.einsum('ask,kb->asb', q, r) - t).max() < 1e-13 * np.abs(t).max()
    q0, v0 = orc.decompose_l2r(t)[:2]                  # v0 = R^T
    assert np.abs(r - v0.T).max() < 1e-11 * np.abs(v0).max() and np.abs(q - q0).max() < 1e-10
    Q, R = be.qr_tensor(be.from_numpy(t), False)
    q, r = be.to_numpy(Q), be.to_numpy(R)
    k = min(a, d * b)
    assert q.shape == (k, d, b) and r.shape == (k, a)
    assert np.abs(np.einsum('xsb,ysb->xy', q, q) - np.eye(k)).max() < 1e-13
    assert np.abs(np.einsum('ka,ksb->asb', r, q) - t).max() < 1e-13 * np.abs(t).max()
    q0, v0 = orc.decompose_r2l(t)[:2]
    assert np.abs(r - v0.T).max() < 1e-11 * np.abs(v0).max() and np.abs(q - q0).max() < 1e-10


def test_no_library_factorisation_on_the_sweep_path(be):
    """the gauge moves of a sweep go through tn_qr_*: torch.linalg.qr (cuSOLVER) must not be called"""
    import torch
    from tnalg_b200 import Parameters as Pm
    from tnalg_b200.DMRG_anyH import sweep_once
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=8, chi=12)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(0)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    calls = []
    real = torch.linalg.qr
    torch.linalg.qr = lambda *a, **k: calls.append(1) or real(*a, **k)
    try:
        A.correct_orthogonal_center(0)
        sweep_once(A, para)
        A.calculate_entanglement_spectrum()
    finally:
        torch.linalg.qr = real
    assert not calls


@pytest.mark.parametrize('n', [1, 2, 7, 64, 200, 512])
def test_jacobi_eigh_vs_lapack(be, n):
    rng = np.random.RandomState(n)
    g = rng.randn(n, n)
    a = (g + g.T) / 2
    w, V = be.eigh(be.from_numpy(a))
    w, V = be.to_numpy(w), be.to_numpy(V)
    w0 = np.linalg.eigvalsh(a)
    assert np.abs(w - w0).max() < 1e-12 * max(1.0, np.abs(w0).max())
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12
    assert np.abs(a @ V - V * w[None, :]).max() < 1e-11 * max(1.0, np.abs(w0).max())


def test_eig_way_0_solves_the_dense_matrix(be, golden):
    """eig_way = 0 (MPSClass.py:792-794): the explicit 1 - tau*H_eff and its dominant eigenvector (Jacobi eigh), same converged
    energies as the operator path and as the reference golden"""
    from tests.test_gpu_dmrg import para_from_golden
    from tnalg_b200.DMRG_anyH import dmrg_finite_size
    g = golden('e2e_chain12')
    from tnalg_b200 import Parameters as Pm

    def make(way):
        para = Pm.generate_parameters_dmrg('chain')
        para.update(l=8, chi=16, eigs_tol=1e-12, break_tol=1e-13, dt_ob=1, eigWay=way)
        return Pm.make_consistent_parameter_dmrg(para)
    para, para2 = make(0), make(1)
    assert g['e_per_site'].size == 1
    np.random.seed(3)
    ob0, A0, info0, _ = dmrg_finite_size(para)
    np.random.seed(3)
    ob1, A1, info1, _ = dmrg_finite_size(para2)
    e_ed = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0] / para['l']
    assert abs(ob0['e_per_site'][0] - ob1['e_per_site'][0]) < 1e-10 and abs(ob0['e_per_site'][0] - e_ed) < 1e-10
    assert info0['n_matvec'] > info1['n_matvec']       # the dense path applies the plan to every unit vector


def test_exact_diagonalisation_on_device_vs_dense(be):
    """f4: EDbasic.project_all_hamilt and the ground state on the full 2^L space against the dense Hamiltonian of the oracle"""
    from tnalg_b200.EDspinClass import EDbasic
    para = orc.make_para('chain', l=10, chi=8, jxy=1.0, jz=0.5, hx=0.0, hz=0.0)
    H = orc.dense_hamiltonian(para)
    _, sx, sy, sz, _, _ = orc.spin_operators('half')
    h2 = np.real(1.0 * (np.kron(sx, sx) + np.kron(sy, sy)) + 0.5 * np.kron(sz, sz))
    L = 10
    couplings = np.array([[i, i + 1, 0] for i in range(L - 1)])
    np.random.seed(0)
    a = EDbasic([2] * L)
    v = np.random.RandomState(1).randn(2 ** L)
    out = a.project_all_hamilt(v, [h2], 1e-3, couplings)
    assert np.abs(out - (v - 1e-3 * H @ v)).max() < 1e-13 * np.abs(v).max()
    e0, vec = a.ground_state([h2], couplings, tau=1e-4, tol=1e-12)
    w = np.linalg.eigvalsh(H)
    assert abs(e0 - w[0]) < 1e-9 * abs(w[0]) and abs(vec @ H @ vec - w[0]) < 1e-10 * abs(w[0])
    # long-range coupling and a second Hamiltonian (sites not adjacent, p1 > p2)
    h3 = np.real(np.kron(sz, sx) + 0.3 * np.kron(sx, sz))
    c2 = np.array([[0, 7, 0], [8, 2, 1], [3, 4, 1]])
    out = a.project_all_hamilt(v, [h2, h3], -0.7, c2)
    t = v.reshape([2] * L)
    ref = t.copy()
    for p1, p2, c in c2:
        hh = [h2, h3][c].reshape(2, 2, 2, 2)
        ref = ref + 0.7 * np.moveaxis(np.tensordot(hh, t, ([2, 3], [p1, p2])), [0, 1], [p1, p2])
    assert np.abs(out - ref.reshape(-1)).max() < 1e-13 * np.abs(ref).max()


def test_eigs_fh_contract(be):
    """a8': Eigs_Module_sjr.eigs_fh(lin_map, d, n, k, v0, tol, max_it, which) with a host callable, a device callable and an
    EffHPlan; n > 1 eigenpairs by deflation; 'sa' / 'la' / 'lm' against dense eigh"""
    import torch
    from tnalg_b200.Eigs_Module_sjr import eigs_fh
    rng = np.random.RandomState(2)
    d = 300
    g = rng.randn(d, d)
    H = (g + g.T) / 2 + np.diag(np.linspace(-20, 25, d))
    w, V = np.linalg.eigh(H)
    Hd = be.from_numpy(H)
    calls = {'host': 0, 'dev': 0}

    def host_map(v):                         # numpy in spirit: the reference's lin_map works on (d, 1) arrays
        calls['host'] += 1
        return H @ v.cpu().numpy()

    def dev_map(v):
        calls['dev'] += 1
        return Hd @ v                         # the caller's own operator, not part of the product path

    lm, v, info = eigs_fh(host_map, d, n=1, k=30, v0=rng.randn(d, 1), tol=1e-12, which='sa')
    assert lm.shape == (1,) and v.shape == (d, 1) and calls['host'] > 0 and info['converged']
    assert abs(lm[0] - w[0]) < 1e-9 * abs(w[0]) and abs(abs(v[:, 0] @ V[:, 0]) - 1) < 1e-8
    lm, v, info = eigs_fh(dev_map, d, n=3, k=40, tol=1e-12, which='sa')
    assert np.abs(lm - w[:3]).max() < 1e-8 * np.abs(w).max()
    for i in range(3):
        assert abs(abs(v[:, i] @ V[:, i]) - 1) < 1e-6
    assert np.abs(v.T @ v - np.eye(3)).max() < 1e-10 and info['error'].shape == (1, 3)
    lm, v, info = eigs_fh(dev_map, d, n=2, k=40, tol=1e-12, which='la')
    assert np.abs(lm - w[::-1][:2]).max() < 1e-8 * np.abs(w).max()
    lm, v, info = eigs_fh(dev_map, d, n=1, k=40, tol=1e-12, which='lm')
    big = w[np.argmax(np.abs(w))]
    assert abs(lm[0] - big) < 1e-8 * abs(big)
    # an effective-Hamiltonian plan as lin_map
    from tnalg_b200 import Parameters as Pm
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=8, chi=8)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(1)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(3)
    plan = A.effective_hamiltonian_plan(3, para['index1'], para['index2'], para['coeff1'], para['coeff2'], tol=1e-12)
    n = int(np.prod(plan.shape))
    hd = A.effective_hamiltonian_dmrg(3, para['index1'], para['index2'], para['coeff1'], para['coeff2'])
    wd = np.linalg.eigvalsh(hd)
    lm, v, info = eigs_fh(plan, n, n=2, tol=1e-12, which='sa')
    assert np.abs(lm - wd[:2]).max() < 1e-9 * np.abs(wd).max()
    with pytest.raises(ValueError):
        eigs_fh(plan, n + 1, n=1)
    with pytest.raises(TypeError):
        eigs_fh(3.0, n)
    plan.destroy()
    assert isinstance(torch.zeros(1), torch.Tensor)


def test_tensor_basic_module_wrappers_on_gpu(be, golden):
    """a6: absorb_matrices2tensor(_full_fast), bound_vec_with_phys_* and friends on the CUDA backend against outputs of the
    reference's own functions (tests/golden/tensor_kats.npz)"""
    from tests.test_host_logic_cpu import check_tensor_helpers
    check_tensor_helpers(golden, be)


def test_observables_chi512_vs_oracle(be):
    """a10 at the bond dimension BASELINE configs[2] names: one- and two-body expectation values on a random centre-orthogonal
    MPS with chi = 512 (L = 22: bonds 1,2,...,512,...,2,1) against the oracle's transfer chains, abs 1e-8 (measured ~1e-13)"""
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    L, d, chi = 22, 2, 512
    oplist = [np.real(o) if np.abs(np.imag(o)).max() == 0 else o for o in orc.spin_operators('half')]
    np.random.seed(11)
    A = MpsOpenBoundaryClass(L, d, chi, operators=oplist)
    A.correct_orthogonal_center(L // 2)
    A.mps[L // 2] = A.mps[L // 2] / A.norm_mps()
    assert max(A.virtual_dim) == 512
    host = [be.to_numpy(t) for t in A.mps]
    O = orc.OracleMps(L, d, chi, oplist, mps=host)
    O.center = L // 2
    mz = A.observe_magnetization(3).reshape(-1)
    mx = A.observe_magnetization(1).reshape(-1)
    for i in (0, 5, 10, 11, 12, 17, 21):
        assert abs(mz[i] - O.observe_one_body(3, i)) < 1e-8 and abs(mx[i] - O.observe_one_body(1, i)) < 1e-8
    index2 = np.array([[9, 10, 4, 5], [10, 11, 3, 3], [11, 12, 5, 4], [3, 18, 3, 3], [10, 13, 1, 1], [0, 21, 3, 3]])
    coeff2 = np.array([0.5, 1.0, 0.5, 0.25, 2.0, 1.0])
    eb = A.observe_bond_energy(index2, coeff2).reshape(-1)
    for r, c, got in zip(index2, coeff2, eb):
        assert abs(got - c * O.observe_two_body([r[2], r[3]], [r[0], r[1]])) < 1e-8
    corr = A.observe_correlators_from_middle(3, 3)
    ref = O.observe_correlators_from_middle(3, 3)
    assert np.abs(corr - ref).max() < 1e-8
    # the one-pass form of the sweep driver, with a (S+ S-) / (S- S+) pair that is contracted once
    index2b = np.array([[9, 10, 4, 5], [9, 10, 5, 4], [10, 11, 3, 3], [3, 18, 3, 3], [2, 7, 4, 5], [2, 7, 5, 4]])
    coeff2b = np.array([0.5, 0.5, 1.0, 0.25, 0.5, 0.5])
    eb2, (mx2, mz2) = A.observe_bond_energy_and_magnetization(index2b, coeff2b, (1, 3))
    for r, c, got in zip(index2b, coeff2b, eb2.reshape(-1)):
        assert abs(got - c * O.observe_two_body([r[2], r[3]], [r[0], r[1]])) < 1e-8
    assert np.abs(mx2.reshape(-1) - mx).max() < 1e-12 and np.abs(mz2.reshape(-1) - mz).max() < 1e-12


@pytest.mark.parametrize('world', [2, 3, 8])
def test_row_sliced_plans_tile_the_full_matvec(be, world):
    """e: the row-sliced plans of `world` ranks (tn_effh_plan_create_rows), run one after the other on one GPU, reproduce the rows
    of the full matvec (bit-level agreement is not required: different tile schedules); uneven slices included"""
    import torch
    from tnalg_b200 import Parameters as Pm
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = Pm.generate_parameters_dmrg('square')
    para.update(square_width=4, square_height=3, chi=100, op=para['op'][:6])
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(2)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    p = 5
    A.correct_orthogonal_center(p)
    args = (para['index1'], para['index2'], para['coeff1'], para['coeff2'])
    full = A.effective_hamiltonian_plan(p, *args, tol=1e-8)
    a, d, b = A.mps[p].shape
    x = be.from_numpy(np.random.RandomState(0).randn(a, d, b))
    want = full.matvec(x, 0.7, -0.3).clone()
    rows_per = -(-a // world)
    got = torch.empty_like(want)
    for r in range(world):
        rb = r * rows_per
        rc = min(rows_per, a - rb)
        if rc <= 0:
            continue
        pr = A.effective_hamiltonian_plan(p, *args, tol=1e-8, rows=(rb, rc))
        assert pr.rows and (pr.row_begin, pr.row_count) == (rb, rc)
        got[rb:rb + rc] = pr.matvec(x, 0.7, -0.3)
        pr.destroy()
    full.destroy()
    assert float((got - want).abs().max()) <= 1e-13 * float(want.abs().max())


def test_deterministic_mode_is_bit_reproducible(be):
    """tn_set_deterministic(1): no FP64-atomic combination of partial tiles -> identical bits from run to run, also with more
    right-stage links than one chunk (the case that needed atomics in round 1)"""
    from tnalg_b200 import Parameters as Pm
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = Pm.generate_parameters_dmrg('square')
    para.update(square_width=4, square_height=4, chi=128, op=para['op'][:6])
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(6)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(7)
    args = (para['index1'], para['index2'], para['coeff1'], para['coeff2'])
    x = be.from_numpy(np.random.RandomState(0).randn(*A.mps[7].shape))
    ref = A.effective_hamiltonian_plan(7, *args, tol=1e-8)
    want = ref.matvec(x).clone()
    ref.destroy()
    old = be.set_deterministic(True)
    try:
        outs = []
        for _ in range(3):
            plan = A.effective_hamiltonian_plan(7, *args, tol=1e-8)
            outs.append(plan.matvec(x).clone())
            outs.append(plan.matvec(x).clone())
            plan.destroy()
    finally:
        be.set_deterministic(old)
    for o in outs[1:]:
        assert bool((o == outs[0]).all())
    assert float((outs[0] - want).abs().max()) <= 1e-13 * float(want.abs().max())


def test_spin1_two_site_sweep_reaches_ed(be):
    """f1 with d*d = 9 (two spin-1 sites): the site operators of the window are applied by one element-wise pass each before the
    GEMMs (TN_MAX_LOADPATH_DIM = 4 < 9); the sweep must reach the exact ground-state energy when chi is exact"""
    from tnalg_b200 import Parameters as Pm
    from tnalg_b200.DMRG_anyH import dmrg_finite_size_two_site
    para = Pm.generate_parameters_dmrg('chain')
    para.update(spin='one', l=6, chi=27, eigs_tol=1e-13, break_tol=1e-13, sweep_time=8, dt_ob=1)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(5)
    ob, A, info, para = dmrg_finite_size_two_site(para, chi_init=3)
    e_ed = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert abs(float(np.ravel(ob['e_per_site'])[0]) * para['l'] - e_ed) < 1e-9 * abs(e_ed)
    assert info['not_converged'] == 0


def test_stale_environment_after_in_place_coupling_edit(be):
    """the environment cache is keyed on the CONTENT of the coupling arrays -- an in-place edit of para['coeff2']
    must give the same energy as a fresh run with the edited couplings"""
    from tnalg_b200 import Parameters as Pm
    from tnalg_b200.DMRG_anyH import observe, sweep_once
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=6, chi=8, eigs_tol=1e-12)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(0)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(0)
    for _ in range(4):
        sweep_once(A, para)
    para['coeff2'][2::3] *= 3.0             # quench J_z in place
    for _ in range(8):
        sweep_once(A, para)
    e_new = observe(A, para, {})['e_per_site'][0] * para['l']
    e_ed = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert abs(e_new - e_ed) < 1e-9 * abs(e_ed)


def test_complex_operator_observable_is_rejected(be):
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    np.random.seed(0)
    A = MpsOpenBoundaryClass(4, 2, 4)
    A.correct_orthogonal_center(1)
    with pytest.raises(ValueError):
        A.observe_magnetization(2)          # sy


def test_library_collectives_on_two_gpus():
    """e: tests/multigpu_check.py under torchrun (NCCL, world 2) when the box has two GPUs"""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29517', os.path.join(root, 'tests', 'multigpu_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and 'multigpu_check ok' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_huge_norms_do_not_overflow(be):
    """un-normalised random MPS tensors of a long chain reach norms of 1e300 (a factor sqrt(d*chi) per site of the initial gauge
    pass); the reference survives because LAPACK/ARPACK norms are scaled -- QR and the Lanczos start vector must as well"""
    rng = np.random.RandomState(0)
    a = rng.randn(300, 120) * 1e298
    Q, R = be.qr(be.from_numpy(a))
    q, r = be.to_numpy(Q), be.to_numpy(R)
    assert np.isfinite(q).all() and np.isfinite(r).all()
    assert np.abs(q.T @ q - np.eye(120)).max() < 1e-13 and np.abs(q @ (r / 1e298) - a / 1e298).max() < 1e-12
    tiny = rng.randn(64, 64) * 1e-300
    Q, R = be.qr(be.from_numpy(tiny))
    q, r = be.to_numpy(Q), be.to_numpy(R)
    assert np.abs(q.T @ q - np.eye(64)).max() < 1e-13 and np.abs(q @ (r * 1e300) - tiny * 1e300).max() < 1e-12
    from tnalg_b200 import Parameters as Pm
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = Pm.generate_parameters_dmrg('chain')
    para.update(l=6, chi=8)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(0)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(2)
    A.mps[2] = A.mps[2] * 1e299
    A.update_tensor_eigs(2, para['index1'], para['index2'], para['coeff1'], para['coeff2'], para['tau'], True, tol=1e-10)
    assert A.last_eig['converged'] and abs(A.norm_mps() - 1) < 1e-12


def test_idmrg_white_two_site_vs_reference_gpu(be, golden):
    """f1 + f3 on the CUDA path: the reference's two-site handle / dense two-site H_eff per call, and the White-style iDMRG
    trajectory (tests/test_idmrg_tebd_cpu.py::check_idmrg, goldens from the unmodified reference)"""
    from tests.test_idmrg_tebd_cpu import check_idmrg
    check_idmrg(golden, be)


def test_tebd_standard_vs_reference_gpu(be, golden):
    """f3 on the CUDA path: tebd_standard against the reference's evolved state and observables"""
    from tests.test_idmrg_tebd_cpu import check_tebd
    check_tebd(golden, be)


def test_expect_c_abi_vs_oracle(be):
    """a10 through tn_expect_1body / tn_expect_2body against the oracle's transfer chains and against the batched
    Python path, centre at the left end, in the middle and at the right end"""
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    L, d, chi = 9, 2, 24
    ops_ = [np.real(o) if np.abs(np.imag(o)).max() == 0 else o for o in orc.spin_operators('half')]
    np.random.seed(21)
    A = MpsOpenBoundaryClass(L, d, chi, operators=ops_)
    for c in (0, 4, 8):
        A.correct_orthogonal_center(c)
        A.mps[c] = A.mps[c] / A.norm_mps()
        O = orc.OracleMps(L, d, chi, ops_, mps=[be.to_numpy(t) for t in A.mps])
        O.center = c
        one = [((i, ops_[3]),) for i in range(L)] + [((2, ops_[1]),), ((7, ops_[4] + ops_[5]),)]
        two = [((0, ops_[3]), (8, ops_[3])), ((3, ops_[4]), (4, ops_[5])), ((1, ops_[1]), (6, ops_[3])), ((5, ops_[3]), (6, ops_[3]))]
        got = be.expect_terms(list(A.mps), c, one + two)
        want = [O.observe_one_body(3, i) for i in range(L)] + [O.observe_one_body(1, 2), O.observe_one_body(4, 7) + O.observe_one_body(5, 7)]
        want += [O.observe_two_body([3, 3], [0, 8]), O.observe_two_body([4, 5], [3, 4]), O.observe_two_body([1, 3], [1, 6]),
                 O.observe_two_body([3, 3], [5, 6])]
        assert np.abs(got - np.real(want)).max() < 1e-12
        assert np.abs(got[:L] - A.observe_magnetization(3).reshape(-1)).max() < 1e-12
