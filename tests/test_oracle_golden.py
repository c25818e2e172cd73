"""Pin oracle/dmrg_oracle.py against vectors produced by the unmodified reference
(tests/golden/make_golden.py) -- CPU only."""
import numpy as np
import pytest

from oracle import dmrg_oracle as orc


def para_from_golden(g, **kw):
    ops = [np.real(o) if np.abs(np.imag(o)).max() == 0 else o for o in g['op']]
    para = orc.make_para('arbitrary', index1=g['index1'], coeff1=g['coeff1'], index2=g['index2'],
                         coeff2=g['coeff2'], op=ops, chi=int(g['chi']), tau=float(g['tau']),
                         eigs_tol=float(g['eigs_tol']), break_tol=float(g['break_tol']),
                         hx=float(g['hx']), hz=float(g['hz']), **kw)
    assert para['l'] == int(g['l'])
    if 'positions_h2' in g:
        para['positions_h2'] = np.asarray(g['positions_h2'])   # bond order of the generator that made the golden
    return para


def test_docstring_known_answers(golden):
    g = golden('docstring_kats')
    # integers printed in TensorBasicModule.py:395-402, :542-550, :588-596
    assert np.array_equal(g['absorb_out'], [[[5, 11], [8, 18]], [[11, 25], [14, 32]]])
    assert np.array_equal(g['l2r_id'], [[23, 14, 12], [14, 19, 9], [12, 9, 11]])
    assert np.array_equal(g['l2r_v'], [[81, 65, 58], [60, 64, 45], [46, 40, 33]])
    assert np.array_equal(g['r2l_id'], [[15, 11, 13], [11, 19, 15], [13, 15, 19]])
    assert np.array_equal(g['r2l_v'], [[55, 54, 57], [53, 58, 59], [48, 51, 50]])
    assert np.array_equal(orc.mode_product(g['absorb_T'], g['absorb_M'], 2), g['absorb_out'])
    assert np.array_equal(orc.transfer_l2r(g['T3']), g['l2r_id'])
    assert np.array_equal(orc.transfer_l2r(g['T3'], env=g['v']), g['l2r_v'])
    assert np.array_equal(orc.transfer_r2l(g['T3']), g['r2l_id'])
    assert np.array_equal(orc.transfer_r2l(g['T3'], env=g['v']), g['r2l_v'])


def test_model_tables_match_reference(golden):
    g = golden('e2e_chain12')
    p = orc.make_para('chain', l=12, chi=16)
    for k in ('index1', 'index2', 'coeff1', 'coeff2', 'positions_h2'):
        assert np.array_equal(p[k], g[k]), k
    assert all(np.allclose(a, b) for a, b in zip(p['op'], g['op']))
    g = golden('e2e_xxz10')
    p = orc.make_para('chain', l=10, chi=12, jxy=1, jz=0.5, hx=0.3, hz=0)
    for k in ('index1', 'index2', 'coeff1', 'coeff2'):
        assert np.array_equal(p[k], g[k]), k
    assert np.allclose(p['op'][6], g['op'][6])
    g = golden('e2e_j1j2_4x2')
    p = orc.j1j2_square_para(4, 2, chi=16)
    for k in ('index1', 'index2', 'coeff1', 'coeff2', 'positions_h2'):
        assert np.array_equal(p[k], g[k]), k


def test_percall_primitives(golden):
    g = golden('percall_j1j2')
    T, E, F, o = g['tr_T'], g['tr_E'], g['tr_F'], g['tr_op']
    assert np.allclose(orc.transfer_l2r(T, o, E), g['tr_l2r'], rtol=0, atol=1e-13)
    assert np.allclose(orc.transfer_l2r(T), g['tr_l2r_id'], rtol=0, atol=1e-13)
    assert np.allclose(orc.transfer_r2l(T, o, F), g['tr_r2l'], rtol=0, atol=1e-13)
    assert np.allclose(orc.transfer_r2l(T), g['tr_r2l_id'], rtol=0, atol=1e-13)
    assert np.allclose(orc.mode_product(T, E, 0), g['mp0'], atol=1e-13)
    assert np.allclose(orc.mode_product(T, o, 1), g['mp1'], atol=1e-13)
    assert np.allclose(orc.mode_product(T, F, 2), g['mp2'], atol=1e-13)
    X = g['dec_X']
    for way in ('qr', 'svd'):
        q, v, k, lm = orc.decompose_l2r(X, way)
        assert np.allclose(np.einsum('asb,kb->ask', q, v), X, atol=1e-13)
        assert np.allclose(np.abs(q), np.abs(g['dec_l2r_%s_q' % way]), atol=1e-12)
        assert np.allclose(lm, g['dec_l2r_%s_lm' % way], atol=1e-13)
        q, v, k, lm = orc.decompose_r2l(X, way)
        assert np.allclose(np.einsum('ka,ksb->asb', v.T, q), X, atol=1e-13)
        assert np.allclose(np.abs(q), np.abs(g['dec_r2l_%s_q' % way]), atol=1e-12)
        assert np.allclose(lm, g['dec_r2l_%s_lm' % way], atol=1e-13)
    assert abs(orc.entanglement_entropy(np.array([2, 1, 0.5, 0.3, 0])) - float(g['ent_kat'])) < 1e-14
    assert abs(float(g['ent_kat']) + 4.981888749420921) < 1e-12   # TensorBasicModule.py:792-794


@pytest.mark.parametrize('p', [0, 2, 4, 5, 8])
def test_percall_matvec_and_dense_heff(golden, p):
    g = golden('percall_j1j2')
    L, d, chi = int(g['l']), int(g['d']), int(g['chi'])
    ops = [np.real(o) for o in g['op']]
    A = orc.OracleMps(L, d, chi, ops, mps=[g['p%d_mps_%d' % (p, n)] for n in range(L)])
    A.center = p
    env = A.grouped_environments(p, g['index1'], g['index2'], g['coeff1'], g['coeff2'], 1e-12)
    assert sorted(env.keys()) == sorted(str(k) for k in g['p%d_keys' % p])
    assert len(env.get('1_0_1', [])) == int(g['p%d_ncross' % p])
    y = A.apply_handle(g['p%d_x' % p], env, tuple(g['p%d_shape' % p]), float(g['tau']))
    assert np.abs(y - g['p%d_y' % p]).max() < 1e-13 * max(1, np.abs(g['p%d_y' % p]).max())
    h = A.dense_effective_hamiltonian(p, g['index1'], g['index2'], g['coeff1'], g['coeff2'])
    assert np.abs(h - g['p%d_heff' % p]).max() < 1e-13
    # handle == dense operator (eig_way 1 vs 0, MPSClass.py:792-803)
    x = g['p%d_x' % p]
    assert np.abs((x - float(g['tau']) * h @ x) - y).max() < 1e-12


def test_percall_observables(golden):
    g = golden('percall_j1j2')
    L, d, chi = int(g['l']), int(g['d']), int(g['chi'])
    A = orc.OracleMps(L, d, chi, [np.real(o) for o in g['op']], mps=[g['ob_mps_%d' % n] for n in range(L)])
    A.center = 4
    assert np.abs(A.observe_magnetization(1) - g['ob_mx']).max() < 1e-13
    assert np.abs(A.observe_magnetization(3) - g['ob_mz']).max() < 1e-13
    assert np.abs(A.observe_bond_energy(g['index2'], g['coeff2']) - g['ob_eb_full']).max() < 1e-13
    assert np.abs(A.observe_correlators_from_middle(3, 3) - g['ob_corr_z']).max() < 1e-13
    assert np.abs(A.observe_correlators_from_middle(1, 1) - g['ob_corr_x']).max() < 1e-13
    assert abs(A.norm() - float(g['ob_norm'])) < 1e-13


@pytest.mark.parametrize('case', ['e2e_chain12', 'e2e_xxz10', 'e2e_j1j2_4x2', 'e2e_spin1_chain8', 'e2e_longrange8', 'e2e_square3x2', 'e2e_periodic8'])
def test_end_to_end_against_reference(golden, case):
    """converged tight-tolerance runs: e_per_site / spectrum rel 1e-10, observables abs 1e-8
    (BASELINE.json north_star tolerances)."""
    g = golden(case)
    para = para_from_golden(g)
    ob, A, info = orc.dmrg_finite_size(para, seed=int(g['seed']))
    assert abs(ob['e_per_site'][0] - g['e_per_site'][0]) <= 1e-10 * abs(g['e_per_site'][0])
    for k in ('eb_full', 'eb', 'mx', 'mz', 'corr_x', 'corr_z'):
        assert np.abs(np.asarray(ob[k]).reshape(-1) - g[k].reshape(-1)).max() < 1e-8, k
    assert np.abs(A.ent - g['ent']).max() < 1e-8
    for n in range(para['l'] - 1):
        ref = g['lm_%d' % n]
        assert np.abs(A.lm[n] - ref).max() <= 1e-10 * ref.max() + 1e-12, n
    assert np.array_equal(A.virtual_dim, g['virtual_dim'])


def test_reference_pr_fixture_chi16(golden):
    """the reference's own result pickle (old coeff2==1 convention): re-run with the pickled para.
    Fixture was produced with eigs_tol=1e-3/break_tol=1e-9 => compare at the self-reproducibility level
    measured in SURVEY.md section 4."""
    g = golden('pr_fixtures')
    t = 'chi16'
    ops = [np.real(o) if np.abs(np.imag(o)).max() == 0 else o for o in g[t + '_para_op']]
    para = orc.make_para('arbitrary', index1=g[t + '_para_index1'], coeff1=g[t + '_para_coeff1'],
                         index2=g[t + '_para_index2'], coeff2=g[t + '_para_coeff2'], op=ops,
                         chi=int(g[t + '_para_chi']), tau=float(g[t + '_para_tau']),
                         eigs_tol=float(g[t + '_para_eigs_tol']), break_tol=float(g[t + '_para_break_tol']),
                         sweep_time=int(g[t + '_para_sweep_time']), dt_ob=int(g[t + '_para_dt_ob']),
                         hx=float(g[t + '_para_hx']), hz=float(g[t + '_para_hz']))
    ob, A, info = orc.dmrg_finite_size(para, seed=0)
    assert abs(ob['e_per_site'][0] - (-0.7230247115155729)) < 1e-10
    assert abs(ob['e_per_site'][0] - g[t + '_e_per_site'].reshape(-1)[0]) < 1e-10
    assert np.abs(ob['eb_full'].reshape(-1) - g[t + '_eb_full'].reshape(-1)).max() < 1e-8
    assert np.abs(A.ent.reshape(-1) - g[t + '_ent'].reshape(-1)).max() < 1e-7


def test_dense_ed_cross_check():
    """independent known answer: DMRG (chi = full) vs dense ED on a small XXZ + field chain."""
    para = orc.make_para('chain', l=8, chi=16, jxy=1, jz=0.5, hx=0.3, hz=0.1, eigs_tol=1e-12, break_tol=1e-13)
    ob, A, info = orc.dmrg_finite_size(para, seed=5)
    e0 = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert abs(ob['e_per_site'][0] * para['l'] - e0) < 1e-10


def test_truncation_against_library_generation(golden):
    """a12 oracle pinned against the reference's library/ generation (library/MPSClass.py:186-247, 909-923)"""
    g = golden('truncation_lib')
    L = int(g['l'])
    ref, lms = orc.truncate_mps([g['mps_in_%d' % n] for n in range(L)], int(g['chi1']))
    out = [g['mps_out_%d' % n] for n in range(L)]
    assert [t.shape for t in ref] == [t.shape for t in out]
    for n in range(L - 1):
        assert np.abs(lms[n] - g['lm_%d' % n]).max() <= 1e-12 * g['lm_%d' % n].max()
    assert abs(abs(orc.mps_overlap(ref, out)) - 1) < 1e-12
