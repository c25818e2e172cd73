"""(f)1 / (f)3: the White-style two-site iDMRG and the standard TEBD drivers against goldens produced by the UNMODIFIED reference
(library/ + algorithms/ generation, tests/golden/make_golden.py idmrg_white_case / tebd_case).  Host logic on the CPU stand-in
here; tests/test_gpu_round2.py runs the same checks on the CUDA path."""
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


def idmrg_para(g):
    from tnalg_b200 import Parameters as Pm
    para = Pm.generate_parameters_infinite_dmrg()
    para.update(dmrg_type='white', jxy=float(g['jxy']), jz=float(g['jz']), hx=float(g['hx']), hz=float(g['hz']), chi=int(g['chi']),
                n_site=2, sweep_time=int(g['sweep_time']), dt_ob=1, break_tol=1e-14, form='center_ort')
    return Pm.make_para_consistent_idmrg(para)


def check_idmrg(golden, be):
    from tnalg_b200.DMRG_anyH import dmrg_infinite_size
    from tnalg_b200.MPSClass import MpsInfinite
    g = golden('idmrg_white_xxz')
    para = idmrg_para(g)
    assert np.array_equal(para['hamilt_index'], g['hamilt_index'])            # parameter tables equal the reference's
    # (f)1 per-call parity: the two-site handle and the dense two-site H_eff of the reference on ITS final blocks
    np.random.seed(5)
    A = MpsInfinite(para['form'], para['d'], para['chi'], 4, n_site=2, dmrg_type='white', hamilt_index=para['hamilt_index'])
    A.bath_op_onsite = be.from_numpy(g['bath'])
    A.effective_ops = [be.from_numpy(m) for m in g['eff_ops']]
    y = A.update_central_tensor_effective_ops_fh(g['psi'], float(g['tau'])).reshape(-1)
    assert np.abs(y - g['handle_out']).max() <= 1e-13 * np.abs(g['handle_out']).max()
    h = A.effective_hamilt_from_op()
    assert np.abs(h - g['heff_dense']).max() <= 1e-12 * np.abs(g['heff_dense']).max()
    # (f)3 the driver: same random start (seed 0, one randn draw), the bond energy after every sweep follows the reference
    np.random.seed(int(g['seed']))
    A, ob, info = dmrg_infinite_size(para)
    ref = g['eb_history']
    assert ob['eb_history'].shape == ref.shape
    # Every step is a deterministic map (dominant eigenvector of a non-degenerate matrix, SVD split), so the bond energies
    # follow the reference sweep by sweep: 24 sweeps to 1e-9 (measured 1e-11).  At sweep 24 the reference's ARPACK call returns
    # the SECOND eigenvector (its start vector is exactly orthogonal to the new ground state, a mirror-parity crossing; the
    # golden value equals the energy of eigenvector #2 of the dense matrix) and its trajectory is an artefact from there on;
    # this implementation keeps following the true dominant eigenvector, which is checked against a dense solve below.
    assert np.abs(ob['eb_history'][:24] - ref[:24]).max() < 1e-9
    assert A.check_orthogonality_mps() and 0 <= 1 - np.linalg.norm(A.lm[0]) < 1e-3     # kept weight (truncated to chi = 10)
    np.random.seed(int(g['seed']))
    B = MpsInfinite(para['form'], para['d'], para['chi'], 4, n_site=2, dmrg_type='white', hamilt_index=para['hamilt_index'])
    B.update_ort_tensor_mps('left'), B.update_bath_onsite(), B.update_effective_ops()
    for t in range(30):
        w, v = np.linalg.eigh(B.effective_hamilt_from_op())
        B.update_central_tensor((para['tau'], 'full'))
        vec = be.to_numpy(B.mps[1]).reshape(-1)
        assert abs(abs(vec @ v[:, 0]) - 1) < 1e-9, t                          # the true ground state of the block at every sweep
        B.update_ort_tensor_mps('left'), B.update_bath_onsite(), B.update_effective_ops()


def tebd_para(g):
    from tnalg_b200 import Parameters as Pm
    para = Pm.generate_parameters_standard_tebd()
    para.update(l=int(g['l']), chi=int(g['chi']), jxy=float(g['jxy']), jz=float(g['jz']), hx=float(g['hx']), hz=float(g['hz']),
                tau0=float(g['tau0']), dtau=float(g['dtau']), taut=int(g['taut']), dt_ob=10, iterate_time=int(g['iterate_time']),
                if_break=True, break_tol=1e-30, save_mode='final')
    return Pm.make_para_consistent_tebd(para)


def check_tebd(golden, be):
    from tnalg_b200.TEBDalgo import split_gate, tebd_standard
    from tnalg_b200.HamiltonianModule import hamiltonian_heisenberg_library
    g = golden('tebd_chain6')
    para = tebd_para(g)
    h = hamiltonian_heisenberg_library('half', 1.0, 1.0, 0.6, 0.125, 0.05)
    g0, g1 = split_gate(h, 0.1, 2)
    u = np.einsum('akb,ckd->acbd', g0, g1).reshape(4, 4)                       # the two halves contract back to exp(-tau h)
    import scipy.linalg as la
    assert np.abs(u - la.expm(-0.1 * h)).max() < 1e-14
    np.random.seed(int(g['seed']))
    mps, ob, para = tebd_standard(para)
    # (the reference's own run records no observation: its loop tests `t == iterate_time`, which range() never yields,
    #  TEBDalgo.py:67; the comparison is on the evolved state itself and on observables evaluated on the reference's final MPS)
    assert int(g['n_obs']) == 0 and len(ob['e_site']) == int(g['taut'])   # here: one observation at the end of every tau stage
    from oracle import dmrg_oracle as orc
    L = int(g['l'])
    ref = [g['mps_%d' % n] for n in range(L)]
    mine = [np.asarray(t) for t in mps.mps]
    ov = abs(orc.mps_overlap(ref, mine)) / np.sqrt(abs(orc.mps_overlap(ref, ref)) * abs(orc.mps_overlap(mine, mine)))
    assert abs(ov - 1) < 1e-9                                                  # same state after 120 gate layers (gauge invariant)
    O = orc.OracleMps(L, 2, 8, [np.asarray(o) for o in para['op'][:6]], mps=[t.copy() for t in ref])
    O.center = int(g['center'])
    nrm = O.norm()
    for i in range(L):
        assert abs(np.ravel(ob['mx'][-1])[i] - O.observe_one_body(1, i) / nrm ** 2) < 1e-8
        assert abs(np.ravel(ob['mz'][-1])[i] - O.observe_one_body(3, i) / nrm ** 2) < 1e-8
    eb_ref = [(1.0 * O.observe_two_body([1, 1], [i, i + 1]) + 1.0 * np.real(O.observe_two_body([2, 2], [i, i + 1]))
               + 0.6 * O.observe_two_body([3, 3], [i, i + 1])) / nrm ** 2 for i in range(L - 1)]
    assert np.abs(np.ravel(ob['eb'][-1]) - np.real(eb_ref)).max() < 1e-8
    assert ob['e_site'][-1] < ob['e_site'][0]                                  # imaginary time lowers the energy from stage to stage


def test_idmrg_white_two_site_vs_reference(golden, cpu_be):
    check_idmrg(golden, cpu_be)


def test_tebd_standard_vs_reference(golden, cpu_be):
    check_tebd(golden, cpu_be)
