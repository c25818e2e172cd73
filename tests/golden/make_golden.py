"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference in-process.

Run in the build container only (needs /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Outputs (committed): tests/golden/*.npz.  Every array comes from reference code (via oracle/ref_shim.py);
nothing here is computed by the oracle or by the CUDA path.

Cases
  e2e_chain12      BASELINE cfg1: Heisenberg open chain N=12 chi=16, tight tolerances, seed 0
  e2e_xxz10        XXZ (jxy=1,jz=0.5) + hx=0.3 chain N=10 chi=12 (one-body terms, '0_s_0' group)
  e2e_longrange8   power-law Ising + transverse field ('longRange' generator, 84 terms, 28 pairs), N=8 chi=16, seed 5
  e2e_square3x2    Heisenberg on the 'square' generator 3x2, chi=8 (exact), seed 6
  e2e_full6        all-to-all Heisenberg ('full' generator) N=6 -- degenerate singlets: ENERGY parity only
  e2e_jigsaw7      spin-1 'jigsaw' chain N=7 chi=27 -- possibly degenerate multiplet: ENERGY parity only
  e2e_periodic8    XXZ + field ring N=8 (bound_cond='periodic': one long bond (0, N-1) on the open MPS), chi=16, seed 9
  e2e_spin1_chain8 spin-1 Heisenberg open chain N=8 chi=12 (d = 3 path, Parameters.py:440-446), seed 3
  e2e_j1j2_4x2     J1-J2 4x2 'arbitrary' lattice chi=16 = exact (crossing '1_0_1' terms; even site count so
                   the ground state is a unique singlet -- 3x3 has a degenerate doublet)
  percall_j1j2     one MPS snapshot + per-call reference outputs: matvec handle, dense H_eff, opt_env key
                   set, transfers, QR/SVD gauge moves, observables
  pr_fixtures      numbers extracted from the reference's own data_dmrg/*.pr result pickles
  docstring_kats   integer known-answer vectors printed in TensorBasicModule.py docstrings
  truncation_lib   library/ generation: truncate_virtual_bonds(chi=5) of a random chi=12 MPS (kept spectrum, tensors)
  idmrg_white_xxz  library/ generation: White-style two-site iDMRG (energy per sweep, Schmidt values, two-site handle per call)
  tebd_chain6      library/ generation: tebd_standard on an XXZ chain in a field, L=6 chi=8 (final MPS, energies, magnetisation)
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim  # noqa: E402

Pm, dm, mc, tm, hm = ref_shim.modules()
TIGHT = dict(eigs_tol=1e-12, break_tol=1e-13)


def pack_run(para, seed):
    """reference run: everything from is_save_op=True except corr_* (stale-cache defect)
    which come from a second run with is_save_op=False."""
    ob, A, info, para = ref_shim.run_finite_dmrg(dict(para), seed)
    p2 = dict(para)
    p2['is_save_op'] = False
    ob2, A2, _, _ = ref_shim.run_finite_dmrg(p2, seed)
    out = {'seed': seed, 'l': para['l'], 'chi': para['chi'], 'd': para['d'], 'tau': para['tau'],
           'eigs_tol': para['eigs_tol'], 'break_tol': para['break_tol'], 'hx': para['hx'], 'hz': para['hz'],
           'index1': np.asarray(para['index1']), 'index2': np.asarray(para['index2']),
           'coeff1': np.asarray(para['coeff1']), 'coeff2': np.asarray(para['coeff2']),
           'positions_h2': np.asarray(para['positions_h2']),
           'op': np.stack([np.asarray(o, dtype=complex) for o in para['op']]),
           'e_per_site': ob['e_per_site'], 'eb_full': ob['eb_full'], 'eb': ob['eb'], 'mx': ob['mx'],
           'mz': ob['mz'], 'corr_x': ob2['corr_x'], 'corr_z': ob2['corr_z'],
           'corr_x_stale': ob['corr_x'], 'corr_z_stale': ob['corr_z'],
           'ent': A.ent, 'virtual_dim': A.virtual_dim, 'convergence': info['convergence'],
           'attrs': np.array(sorted(A.__dict__.keys())), 'mps_dtype': str(A.mps[0].dtype)}
    for n, lm in enumerate(A.lm):
        out['lm_%d' % n] = lm
    return out


def chain_para(**kw):
    para = Pm.generate_parameters_dmrg('chain')
    para.update(kw)
    return Pm.make_consistent_parameter_dmrg(para)


def j1j2_para(w, h, chi, j2=0.5, **kw):
    """the working cfg4 recipe of SURVEY.md sec. 8d on a (w x h) lattice."""
    nn = hm.positions_nearest_neighbor_square(w, h, 'open').astype(int)
    diag = []
    for r in range(h - 1):
        for c in range(w - 1):
            diag.append([r * w + c, (r + 1) * w + c + 1])
            diag.append([r * w + c + 1, (r + 1) * w + c])
    pos = np.vstack([nn, np.array(diag, dtype=int)])
    jj = np.concatenate([np.ones(nn.shape[0]), j2 * np.ones(len(diag))])
    para = Pm.generate_parameters_dmrg('chain')  # common keys; then overwritten as 'arbitrary'
    op = hm.spin_operators('half')
    L = w * h
    para.update(lattice='arbitrary', spin='half', hx=0, hz=0, chi=chi,
                op=[op['id'], op['sx'], op['sy'], op['sz'], op['su'], op['sd'], np.zeros((2, 2))],
                index1=[[i, 6] for i in range(L)], coeff1=np.ones(L),
                index2=hm.interactions_position2full_index_heisenberg_two_body(pos),
                coeff2=np.stack([jj / 2, jj / 2, jj], axis=1).reshape(-1))
    para.update(kw)
    # `is 'arbitrary'` identity comparison in Parameters.py:205 needs the interned literal
    para['lattice'] = sys.intern('arbitrary')
    return Pm.make_consistent_parameter_dmrg(para)


def percall_case():
    """Snapshot of a partially converged reference MPS on the 3x3 J1-J2 lattice and the outputs of the
    reference's own functions on it."""
    para = j1j2_para(3, 3, 8, sweep_time=2, dt_ob=2, **TIGHT)
    ob, A, info, para = ref_shim.run_finite_dmrg(dict(para), 3)
    # rebuild a live object (clean_to_save dropped the caches), cache-free mode
    B = mc.MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], way='qr', ini_way='r', operators=para['op'],
                                is_save_op=False, eig_way=1, is_env_parallel_lmr=False)
    B.mps = [t.copy() for t in A.mps]
    B.virtual_dim = A.virtual_dim.copy()
    B.center = A.center
    B.orthogonality = A.orthogonality.copy()
    out = {'l': para['l'], 'd': para['d'], 'chi': para['chi'], 'center0': A.center,
           'index1': para['index1'], 'index2': para['index2'], 'coeff1': para['coeff1'],
           'coeff2': para['coeff2'], 'op': np.stack([np.asarray(o, dtype=complex) for o in para['op']])}
    for n, t in enumerate(A.mps):
        out['mps_%d' % n] = t
    rng = np.random.RandomState(11)
    tau = 0.37
    for p in (0, 2, 4, 5, 8):
        B.correct_orthogonal_center(p)
        for n, t in enumerate(B.mps):
            out['p%d_mps_%d' % (p, n)] = t.copy()
        B.opt_env = dict()
        s = B.all_environments_optimized(p, para['index1'], para['index2'], para['coeff1'], para['coeff2'],
                                         tol=1e-12)
        keys = sorted(B.opt_env.keys())
        out['p%d_keys' % p] = np.array(keys)
        out['p%d_ncross' % p] = len(B.opt_env.get('1_0_1', []))
        x = rng.randn(int(np.prod(s)))
        out['p%d_x' % p] = x
        out['p%d_y' % p] = B.update_tensor_eigs_f_handle_optimized(x, s, tau)
        out['p%d_shape' % p] = np.array(s)
        B.opt_env = dict()
        h, _ = B.effective_hamiltonian_dmrg(p, para['index1'], para['index2'], para['coeff1'], para['coeff2'])
        out['p%d_heff' % p] = h
    out['tau'] = tau
    # transfers / gauge moves on one tensor
    T = A.mps[4]
    E = rng.randn(T.shape[0], T.shape[0])
    F = rng.randn(T.shape[2], T.shape[2])
    o = np.real(para['op'][4])
    out.update(tr_T=T, tr_E=E, tr_F=F, tr_op=o,
               tr_l2r=tm.bound_vec_operator_left2right(T, o, E), tr_l2r_id=tm.bound_vec_operator_left2right(T),
               tr_r2l=tm.bound_vec_operator_right2left(T, o, F), tr_r2l_id=tm.bound_vec_operator_right2left(T),
               mp0=tm.absorb_matrix2tensor(T, E, 0), mp1=tm.absorb_matrix2tensor(T, o, 1),
               mp2=tm.absorb_matrix2tensor(T, F, 2))
    X = rng.randn(6, 2, 5)
    for way in ('qr', 'svd'):
        q, v, k, lm = tm.left2right_decompose_tensor(X, way)
        out['dec_l2r_%s_q' % way], out['dec_l2r_%s_v' % way], out['dec_l2r_%s_lm' % way] = q, v, lm
        q, v, k, lm = tm.right2left_decompose_tensor(X, way)
        out['dec_r2l_%s_q' % way], out['dec_r2l_%s_v' % way], out['dec_r2l_%s_lm' % way] = q, v, lm
    out['dec_X'] = X
    # observables on the snapshot (cache-free => correct), centre 4
    B.correct_orthogonal_center(4)
    for n, t in enumerate(B.mps):
        out['ob_mps_%d' % n] = t.copy()
    out['ob_mx'] = B.observe_magnetization(1)
    out['ob_mz'] = B.observe_magnetization(3)
    out['ob_eb_full'] = B.observe_bond_energy(para['index2'], para['coeff2'])
    out['ob_corr_z'] = B.observe_correlators_from_middle(3, 3)
    out['ob_corr_x'] = B.observe_correlators_from_middle(1, 1)
    out['ob_norm'] = B.norm_mps()
    out['ent_kat'] = tm.entanglement_entropy(np.array([2, 1, 0.5, 0.3, 0]))
    return out


def pr_fixtures():
    """Numbers held by the reference's own result pickles (data_dmrg/*.pr, old coeff2==1 convention)."""
    out = {}
    for tag in ('chi16', 'chi24'):
        path = os.path.join(ref_shim.REFERENCE_ROOT, 'data_dmrg', 'chainN12_j(1,1)_h(0,0)_%sopen.pr' % tag)
        with open(path, 'rb') as f:
            data = pickle.load(f)
        ob, A, info, para = data['ob'], data['A'], data['info'], data['para']
        for k in ('e_per_site', 'eb_full', 'eb', 'mx', 'mz'):
            out['%s_%s' % (tag, k)] = np.asarray(ob[k])
        out['%s_ent' % tag] = A.ent
        out['%s_virtual_dim' % tag] = A.virtual_dim
        for n, lm in enumerate(A.lm):
            out['%s_lm_%d' % (tag, n)] = lm
        for k in ('l', 'chi', 'd', 'tau', 'eigs_tol', 'break_tol', 'sweep_time', 'dt_ob', 'hx', 'hz',
                  'ob_position'):
            out['%s_para_%s' % (tag, k)] = para[k]
        for k in ('index1', 'index2', 'coeff1', 'coeff2', 'positions_h2'):
            out['%s_para_%s' % (tag, k)] = np.asarray(para[k])
        out['%s_para_op' % tag] = np.stack([np.asarray(o, dtype=complex) for o in para['op']])
        out['%s_t_cost' % tag] = info['t_cost']
        out['%s_attrs' % tag] = np.array(sorted(A.__dict__.keys()))
        out['%s_center' % tag] = int(A.center)
        for n, t in enumerate(A.mps):
            out['%s_mps_%d' % (tag, n)] = np.asarray(t)
    return out


def docstring_kats():
    """Integer examples printed in TensorBasicModule.py docstrings (:395-402, :542-550, :588-596), evaluated
    by the reference code itself."""
    T = np.array([[[1, 2], [2, 3]], [[3, 4], [4, 5]]])
    M = np.array([[1, 3], [2, 4]])
    T3 = np.array([[[1, 2, 1], [2, 1, 2]], [[2, 0, 2], [1, 3, 1]], [[3, 1, 0], [2, 2, 1]]])
    v = np.array([[1, 1, 1], [1, 2, 1], [2, 2, 1]])
    return dict(absorb_T=T, absorb_M=M, absorb_out=tm.absorb_matrix2tensor(T, M, 2),
                T3=T3, v=v, l2r_id=tm.bound_vec_operator_left2right(T3),
                l2r_v=tm.bound_vec_operator_left2right(T3, v=v),
                r2l_id=tm.bound_vec_operator_right2left(T3), r2l_v=tm.bound_vec_operator_right2left(T3, v=v))


def truncation_case():
    """a12: the library/ generation's SVD truncation of every bond (library/MPSClass.py:186-247, 909-923), run from the
    package-style tree (library.MPSClass); needs a stub for matplotlib.cm on top of the shim."""
    import importlib
    import types
    if 'matplotlib.cm' not in sys.modules:
        sys.modules['matplotlib.cm'] = types.ModuleType('matplotlib.cm')
        sys.modules['matplotlib'].cm = sys.modules['matplotlib.cm']
    lib = importlib.import_module('library.MPSClass')
    np.random.seed(3)
    A = lib.MpsOpenBoundaryClass(8, 2, 12, way='qr', ini_way='r')
    out = {'chi1': 5, 'center': 7, 'l': 8, 'd': 2, 'chi0': 12}
    for n, t in enumerate(A.mps):
        out['mps_in_%d' % n] = t.copy()
    A.correct_orthogonal_center(0)
    A.truncate_virtual_bonds(5, center=7, way='full')
    for n, t in enumerate(A.mps):
        out['mps_out_%d' % n] = t.copy()
    for n, lm in enumerate(A.lm):
        out['lm_%d' % n] = lm.copy()
    out['virtual_dim'] = np.asarray(A.virtual_dim)
    return out


def rdm_case():
    """two-body reduced density matrices of a converged reference state (MPSClass.py:841-855), centre left of, between and
    right of the pair; XXZ + field chain N=8, chi=16 (exact), seed 4."""
    para = chain_para(l=8, chi=16, jxy=1, jz=0.7, hx=0.4, hz=0.15, **TIGHT)
    ob, A, info, para = ref_shim.run_finite_dmrg(dict(para), 4)
    out = {'seed': 4, 'l': para['l'], 'chi': para['chi'], 'd': para['d'], 'tau': para['tau'], 'eigs_tol': para['eigs_tol'],
           'break_tol': para['break_tol'], 'hx': para['hx'], 'hz': para['hz'], 'index1': np.asarray(para['index1']),
           'index2': np.asarray(para['index2']), 'coeff1': np.asarray(para['coeff1']), 'coeff2': np.asarray(para['coeff2']),
           'op': np.stack([np.asarray(o, dtype=complex) for o in para['op']]), 'e_per_site': ob['e_per_site']}
    pairs = [(3, 4), (1, 6), (0, 7), (2, 3), (5, 7)]
    out['pairs'] = np.array(pairs)
    for c in (0, 4, 7):
        A.correct_orthogonal_center(c)
        for p1, p2 in pairs:
            out['rdm_c%d_%d_%d' % (c, p1, p2)] = A.reduced_density_matrix_two_body(p1, p2)
    return out


def tensor_kats():
    """outputs of the reference's own tensor helpers next to the hot path (a6 + the physical-bond transfers used by the
    two-body density matrix) on seeded random inputs: TensorBasicModule.py:427-528 (absorb_matrices2tensor[_full_fast]),
    :622-649 (bound_vec_with_phys_*), :652-674 (transfer_matrix_mps), :755-785 (normalize_tensor), :876-900
    (check_orthogonality)."""
    rng = np.random.RandomState(11)
    out = {}
    T = rng.randn(4, 3, 5)
    M = [rng.randn(4, 6), rng.randn(3, 3), rng.randn(5, 2)]
    out['T'] = T
    for i, m in enumerate(M):
        out['M%d' % i] = m
    out['absorb_all'] = tm.absorb_matrices2tensor(T.copy(), [m.copy() for m in M])
    out['absorb_full_fast'] = tm.absorb_matrices2tensor_full_fast(T.copy(), [m.copy() for m in M])
    out['absorb_bonds_2_0'] = tm.absorb_matrices2tensor(T.copy(), [M[2].copy(), M[0].copy()], bonds=[2, 0])
    v2l, v2r = rng.randn(4, 4), rng.randn(5, 5)
    v4l, v4r = rng.randn(3, 3, 4, 4), rng.randn(3, 3, 5, 5)
    out.update(v2l=v2l, v2r=v2r, v4l=v4l, v4r=v4r)
    out['phys_l2r_empty'] = tm.bound_vec_with_phys_left2right(T)
    out['phys_l2r_v2'] = tm.bound_vec_with_phys_left2right(T, v2l)
    out['phys_l2r_v4'] = tm.bound_vec_with_phys_left2right(T, v4l)
    out['phys_r2l_empty'] = tm.bound_vec_with_phys_right2left(T)
    out['phys_r2l_v2'] = tm.bound_vec_with_phys_right2left(T, v2r)
    out['phys_r2l_v4'] = tm.bound_vec_with_phys_right2left(T, v4r)
    out['transfer_matrix'] = tm.transfer_matrix_mps(T)
    tn, nrm = tm.normalize_tensor(T.copy())
    out['normalized'], out['norm'] = tn, nrm
    q = np.linalg.qr(rng.randn(12, 5))[0].reshape(4, 3, 5)
    out['Q'] = q
    out['check_ort_Q_2'] = np.array(bool(tm.check_orthogonality(q, [2], tol=1e-12)))
    out['check_ort_Q_0'] = np.array(bool(tm.check_orthogonality(q, [0], tol=1e-12)))
    out['ones_mps_1'] = tm.ones_open_mps(3, 2, 3)[1]
    return out


def _library_modules():
    """the package-style generation (library/, algorithms/) under the same shim"""
    import importlib
    import types
    if 'matplotlib.cm' not in sys.modules:
        sys.modules['matplotlib.cm'] = types.ModuleType('matplotlib.cm')
        sys.modules['matplotlib'].cm = sys.modules['matplotlib.cm']
    return (importlib.import_module('library.MPSClass'), importlib.import_module('library.Parameters'),
            importlib.import_module('algorithms.DMRG_anyH'), importlib.import_module('algorithms.TEBDalgo'))


def idmrg_white_case():
    """f1 + f3: the reference's White-style TWO-SITE iDMRG (algorithms/DMRG_anyH.py:106-176 driving library/MPSClass.py
    MpsInfinite: term-summed two-site matvec :1688-1707, SVD truncation :1676-1686) on an XXZ chain in a tilted field (no
    symmetry multiplets, so the truncation is unambiguous): the bond energy after every sweep for 40 sweeps from seed 0, the
    final Schmidt values, and per-call vectors of the two-site handle and of the dense effective Hamiltonian on the final blocks."""
    import contextlib
    import io
    import re
    lib, pml, alg, _ = _library_modules()
    para = pml.generate_parameters_infinite_dmrg()
    para.update(dmrg_type=sys.intern('white'), jxy=1, jz=0.5, hx=0.3, hz=0.1, chi=10, n_site=2, sweep_time=40, dt_ob=1,
                break_tol=1e-14, form=sys.intern('center_ort'))
    with contextlib.redirect_stdout(io.StringIO()):
        para = pml.make_para_consistent_idmrg(para)
    np.random.seed(0)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        A, ob, info = alg.dmrg_infinite_size(para)
    eb = np.array([float(x) for x in re.findall(r'Eb = (-?[0-9.e+-]+)', buf.getvalue())])
    assert eb.size == 40
    out = {'seed': 0, 'chi': 10, 'd': 2, 'tau': para['tau'], 'jxy': 1.0, 'jz': 0.5, 'hx': 0.3, 'hz': 0.1, 'sweep_time': 40,
           'hamilt_index': np.asarray(para['hamilt_index'], dtype=float), 'eb_history': eb, 'eb': ob['eb'], 'lm': np.asarray(A.lm[0]),
           'bath': np.real(A.bath_op_onsite), 'eff_ops': np.stack([np.real(np.asarray(o)) for o in A.effective_ops]),
           'op': np.stack([np.asarray(o, dtype=complex) for o in A.operators])}
    psi = np.random.RandomState(7).randn(10, 2, 2, 10)
    out['psi'] = psi
    out['handle_out'] = np.real(A.update_central_tensor_effective_ops_fh(psi.reshape(-1), para['tau'])).reshape(-1)
    out['heff_dense'] = np.real(A.effective_hamilt_from_op())
    return out


def tebd_case():
    """f3: the reference's tebd_standard (algorithms/TEBDalgo.py:11-90; gates :38-46, MpsStandardTEBD library/MPSClass.py:1408-1447)
    on an XXZ chain in a field, L = 6, chi = 8 (exact), two time steps of 60 iterations from seed 2: final energies, magnetisation,
    Schmidt spectrum."""
    import contextlib
    import io
    lib, pml, _, tebd = _library_modules()
    para = pml.generate_parameters_standard_tebd()
    para.update(l=6, chi=8, jxy=1, jz=0.6, hx=0.25, hz=0.1, tau0=0.1, dtau=0.5, taut=2, dt_ob=10, iterate_time=60, if_break=True,
                break_tol=1e-30, save_mode=sys.intern('final'))
    para = pml.make_para_consistent_tebd(para)
    np.random.seed(2)
    with contextlib.redirect_stdout(io.StringIO()):
        mps = lib.MpsStandardTEBD(para['l'], para['d'], para['chi'], para['spin'], evolve_way='gates')
        ini = [t.copy() for t in mps.mps]
    np.random.seed(2)
    with contextlib.redirect_stdout(io.StringIO()):
        mps, ob, para = tebd.tebd_standard(para)
    out = {'seed': 2, 'l': 6, 'chi': 8, 'd': 2, 'jxy': 1.0, 'jz': 0.6, 'hx': 0.25, 'hz': 0.1, 'tau0': 0.1, 'dtau': 0.5, 'taut': 2,
           'iterate_time': 60, 'n_obs': len(ob['e_site'])}
    for n, t in enumerate(ini):
        out['ini_%d' % n] = t
    for n, t in enumerate(mps.mps):
        out['mps_%d' % n] = np.real(t)
    out['center'] = mps.center
    for k in ('e_site', 'mx', 'mz', 'eb'):
        out[k] = np.real(np.asarray(ob[k][-1], dtype=complex)).astype(float) if len(ob[k]) else np.zeros(0)
    return out


def lattice_para(lattice, **kw):
    para = Pm.generate_parameters_dmrg(lattice)
    para.update(kw)
    if lattice == 'square':
        para['op'] = para['op'][:6]        # the generator appends the field operator on every call (SURVEY 8d gotcha ii)
    return Pm.make_consistent_parameter_dmrg(para)


def main():
    cases = {
        'e2e_chain12': lambda: pack_run(chain_para(l=12, chi=16, **TIGHT), 0),
        'e2e_xxz10': lambda: pack_run(chain_para(l=10, chi=12, jxy=1, jz=0.5, hx=0.3, hz=0, **TIGHT), 1),
        'e2e_j1j2_4x2': lambda: pack_run(j1j2_para(4, 2, 16, **TIGHT), 2),
        'e2e_longrange8': lambda: pack_run(lattice_para('longRange', l=8, chi=16, jxy=0, jz=1, hx=0.5, hz=0, alpha=1.0, **TIGHT), 5),
        'e2e_square3x2': lambda: pack_run(lattice_para('square', square_width=3, square_height=2, chi=8, **TIGHT), 6),
        'e2e_full6': lambda: pack_run(lattice_para('full', l=6, chi=8, **TIGHT), 7),
        'e2e_jigsaw7': lambda: pack_run(lattice_para('jigsaw', l=7, chi=27, **TIGHT), 8),
        'e2e_periodic8': lambda: pack_run(chain_para(l=8, chi=16, bound_cond='periodic', jxy=1, jz=0.8, hx=0.2, hz=0, **TIGHT), 9),
        'e2e_spin1_chain8': lambda: pack_run(chain_para(l=8, chi=12, spin=sys.intern('one'), **TIGHT), 3),
        'percall_j1j2': percall_case,
        'pr_fixtures': pr_fixtures,
        'docstring_kats': docstring_kats,
        'truncation_lib': truncation_case,
        'rdm_xxz8': rdm_case,
        'tensor_kats': tensor_kats,
        'idmrg_white_xxz': idmrg_white_case,
        'tebd_chain6': tebd_case,
    }
    only = sys.argv[1:]
    for name, fn in cases.items():
        if only and name not in only:
            continue
        data = fn()
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **data)
        print('%-16s %7.1f KiB  %d arrays' % (name, os.path.getsize(path) / 1024, len(data)))


if __name__ == '__main__':
    main()
