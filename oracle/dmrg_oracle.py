"""CPU oracle for the finite-size DMRG hot path of ranshiju/T-Nalg  --  TEST INFRASTRUCTURE ONLY.

A plain numpy/scipy restatement of the reference algorithm (root-level generation, one-site, fixed chi,
QR gauge) used as the checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; the product package
(tnalg_b200/) never does and fails loudly without its CUDA library.

PARITY PINNING.  This restatement is pinned against the *unmodified* reference executed in the build
container (oracle/ref_shim.py) by tests/golden/make_golden.py; the resulting vectors are committed under
tests/golden/*.npz and re-checked by tests/test_oracle_golden.py (CPU, `-m "not gpu"`).  It is also pinned
against the numbers extracted from the reference's own data_dmrg/*.pr result pickles and against the
integer docstring examples of TensorBasicModule.py.

Third-party arithmetic: the reference's eigensolver is scipy.sparse.linalg.eigsh (ARPACK, unpinned
version; MPSClass.py:4,804).  The oracle calls the same scipy entry point; the eigensolver boundary is
therefore pinned at the level of converged eigenpairs / observables, not ARPACK trajectories.

All citations are file:line relative to /root/reference.
"""
import numpy as np
from scipy.sparse.linalg import LinearOperator, eigsh

# --------------------------------------------------------------------------------------------------
# Model tables (HamiltonianModule.py:8-44, 151-261; Parameters.py:170-281)
# --------------------------------------------------------------------------------------------------


def spin_operators(spin='half'):
    """[id, sx, sy, sz, su, sd] as in HamiltonianModule.py:8-44 (sy complex, never used by the path)."""
    if spin == 'half':
        sx = np.array([[0, 0.5], [0.5, 0]])
        sy = np.array([[0, 0.5j], [-0.5j, 0]])
        sz = np.array([[0.5, 0], [0, -0.5]])
        su = np.array([[0., 1], [0, 0]])
        sd = np.array([[0., 0], [1, 0]])
        return [np.eye(2), sx, sy, sz, su, sd]
    if spin == 'one':
        r = 2 ** 0.5
        sx = np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]]) / r
        sy = np.array([[0, -1j, 0], [1j, 0, -1j], [0, 1j, 0]]) / r
        sz = np.diag([1., 0, -1])
        return [np.eye(3), sx, sy, sz, np.real(sx + 1j * sy), np.real(sx - 1j * sy)]
    raise ValueError(spin)


def heisenberg_index2(positions):
    """3 rows per bond: (su,sd), (sd,su), (sz,sz)  (HamiltonianModule.py:242-261)."""
    rows = []
    for i, j in np.asarray(positions, dtype=int):
        rows += [[i, j, 4, 5], [i, j, 5, 4], [i, j, 3, 3]]
    return np.array(rows, dtype=int).reshape(-1, 4)


def chain_positions(length):
    """open chain nearest neighbours (HamiltonianModule.py:151-163)."""
    return np.array([[n, n + 1] for n in range(length - 1)], dtype=int)


def square_positions(width, height):
    """open square lattice, site = row*width+col; rows first, then columns (HamiltonianModule.py:203-216)."""
    pos = []
    for r in range(height):
        for c in range(width - 1):
            pos.append([r * width + c, r * width + c + 1])
    for c in range(width):
        for r in range(height - 1):
            pos.append([r * width + c, (r + 1) * width + c])
    return np.array(pos, dtype=int)


def make_para(lattice='chain', **kw):
    """Parameter dict with the keys dmrg_finite_size reads (Parameters.py:6-25,64-76,97-111,170-204;
    DMRG_anyH.py:26-91).  lattice in {'chain','square','arbitrary'}; 'arbitrary' needs index1/2, coeff1/2."""
    para = dict(chi=30, sweep_time=100, if_print_detail=False, tau=1e-4, eigs_tol=1e-5, break_tol=1e-8,
                is_real=True, dt_ob=4, ob_position=0, eigWay=1, isParallel=False, isParallelEnvLMR=False,
                is_save_op=True, spin='half', bound_cond='open', jxy=1, jz=1, hx=0, hz=0, lattice=lattice)
    if lattice == 'chain':
        para['l'] = 18
    elif lattice == 'square':
        para.update(square_width=4, square_height=4)
    para.update(kw)
    ops = spin_operators(para['spin'])
    d = ops[0].shape[0]
    if lattice in ('chain', 'square'):
        if lattice == 'square':
            para['l'] = para['square_width'] * para['square_height']
            pos = square_positions(para['square_width'], para['square_height'])
        else:
            pos = chain_positions(para['l'])
        L = para['l']
        para['op'] = ops + [-para['hx'] * ops[1] - para['hz'] * ops[3]]
        para['index1'] = np.stack([np.arange(L), 6 * np.ones(L, dtype=int)], axis=1).astype(int)
        para['coeff1'] = np.ones((L, 1))
        para['positions_h2'] = pos
        para['index2'] = heisenberg_index2(pos)
        para['coeff2'] = np.tile(np.array([para['jxy'] / 2, para['jxy'] / 2, para['jz']]),
                                 pos.shape[0]).reshape(-1, 1).astype(float)
    elif lattice == 'arbitrary':
        para['index1'] = np.array(para['index1'], dtype=int).reshape(-1, 2)
        para['index2'] = np.array(para['index2'], dtype=int).reshape(-1, 4)
        para['coeff1'] = np.array(para['coeff1'], dtype=float).reshape(-1, 1)
        para['coeff2'] = np.array(para['coeff2'], dtype=float).reshape(-1, 1)
        para['l'] = int(max(para['index1'][:, 0].max(), para['index2'][:, :2].max()) + 1)
        para['positions_h2'] = unique_positions(para['index2'])
        para.setdefault('op', ops + [np.zeros((d, d))])
    else:
        raise ValueError(lattice)
    para['d'] = int(np.asarray(para['op'][0]).shape[0])   # user-supplied operators decide ('arbitrary' with spin-1 ops)
    para['nh'] = para['index2'].shape[0]
    return para


def unique_positions(index2):
    """distinct (site1, site2) pairs, sorted (Parameters.py:410-418 with DMRG_anyH.sort_positions)."""
    pairs = sorted({(int(r[0]), int(r[1])) for r in index2})
    return np.array(pairs, dtype=int).reshape(-1, 2)


def j1j2_square_para(width, height, j1=1.0, j2=0.5, **kw):
    """J1-J2 Heisenberg on an open square lattice as a lattice='arbitrary' dict (SURVEY.md 8d recipe:
    NN bonds of positions_nearest_neighbor_square + both diagonals of every plaquette)."""
    nn = square_positions(width, height)
    diag = []
    for r in range(height - 1):
        for c in range(width - 1):
            diag.append([r * width + c, (r + 1) * width + c + 1])
            diag.append([r * width + c + 1, (r + 1) * width + c])
    diag = np.array(diag, dtype=int).reshape(-1, 2)
    pos = np.vstack([nn, diag]) if diag.size else nn
    jj = np.concatenate([np.full(nn.shape[0], j1), np.full(diag.shape[0], j2)])
    index2 = heisenberg_index2(pos)
    coeff2 = np.stack([jj / 2, jj / 2, jj], axis=1).reshape(-1, 1)
    L = width * height
    index1 = np.stack([np.arange(L), 6 * np.ones(L, dtype=int)], axis=1)
    return make_para('arbitrary', index1=index1, coeff1=np.ones(L), index2=index2, coeff2=coeff2, **kw)


# --------------------------------------------------------------------------------------------------
# Tensor primitives (TensorBasicModule.py)
# --------------------------------------------------------------------------------------------------


def random_open_mps(length, d, chi):
    """np.random.randn draw order: site 0, site L-1, sites 1..L-2 (TensorBasicModule.py:181-186)."""
    mps = [None] * length
    mps[0] = np.random.randn(1, d, chi)
    mps[length - 1] = np.random.randn(chi, d, 1)
    for n in range(1, length - 1):
        mps[n] = np.random.randn(chi, d, chi)
    return mps


def mode_product(tensor, mat, bond):
    """out[.., j, ..] = sum_i tensor[.., i, ..] mat[i, j] on axis `bond`
    (absorb_matrix2tensor, TensorBasicModule.py:387-424: the FIRST index of mat is contracted)."""
    out = np.tensordot(tensor, mat, axes=([bond], [0]))
    return np.moveaxis(out, -1, bond)


def transfer_l2r(tensor, op=None, env=None):
    """E'[b,b'] = sum conj(T[a,s,b]) E[a,a'] op[s,s'] T[a',s',b']
    (bound_vec_operator_left2right, TensorBasicModule.py:530-573; env/op None = identity)."""
    a, d, b = tensor.shape
    # the same three matrix products as the reference (:554-568): op on the physical bond, v . T, T^H . (v T)
    ket = tensor if op is None else np.moveaxis(np.tensordot(np.asarray(op), tensor, axes=([1], [1])), 0, 1)
    ket = np.ascontiguousarray(ket)
    if env is not None:
        ket = env.dot(ket.reshape(a, d * b))
    return tensor.conj().reshape(a * d, b).T.dot(ket.reshape(a * d, b))


def transfer_r2l(tensor, op=None, env=None):
    """E'[a,a'] = sum conj(T[a,s,b]) E[b,b'] op[s,s'] T[a',s',b']
    (bound_vec_operator_right2left, TensorBasicModule.py:576-619)."""
    a, d, b = tensor.shape
    ket = tensor if op is None else np.moveaxis(np.tensordot(np.asarray(op), tensor, axes=([1], [1])), 0, 1)
    ket = np.ascontiguousarray(ket)
    bra = tensor.conj().reshape(a * d, b)
    if env is not None:
        bra = bra.dot(env)
    return bra.reshape(a, d * b).dot(ket.reshape(a, d * b).T)


def decompose_l2r(tensor, way='qr'):
    """(a,d,b) -> Q (a,d,k), v = R^T (b,k), k = min(a*d, b), lm
    (left2right_decompose_tensor, TensorBasicModule.py:314-348; note the returned matrix is transposed)."""
    a, d, b = tensor.shape
    k = min(a * d, b)
    mat = tensor.reshape(a * d, b)
    if way in (1, 'svd'):
        u, lm, vh = np.linalg.svd(mat, full_matrices=False)
        r = lm[:k, None] * vh[:k, :]
        q = u
    else:
        q, r = np.linalg.qr(mat)
        lm = np.zeros(0)
    return q[:, :k].reshape(a, d, k), r.T, k, lm


def decompose_r2l(tensor, way='qr'):
    """(a,d,b) -> Q (k,d,b), v = R^T (a,k), k = min(a, d*b), lm
    (right2left_decompose_tensor, TensorBasicModule.py:351-384)."""
    a, d, b = tensor.shape
    k = min(a, d * b)
    mat = tensor.reshape(a, d * b).T
    if way in (1, 'svd'):
        u, lm, vh = np.linalg.svd(mat, full_matrices=False)
        r = lm[:k, None] * vh[:k, :]
        q = u
    else:
        q, r = np.linalg.qr(mat)
        lm = np.zeros(0)
    return q[:, :k].T.reshape(k, d, b), r.T, k, lm


def entanglement_entropy(lm, tol=1e-20):
    """-2 sum lm^2 ln lm over lm > tol (TensorBasicModule.py:786-801)."""
    lm = np.sort(np.asarray(lm).reshape(-1))[::-1]
    lm = lm[lm > tol]
    return float(-2 * np.dot(lm ** 2, np.log(lm)))


def svd_truncate_two_site(theta, chi):
    """theta (a,d,d,b) -> U (a,d,k), lm (k,), Vh (k,d,b), k = min(chi, a*d)
    (library/MPSClass.py:1676-1686, the two-site SVD-truncate; the reference stores Vh transposed to (b,d,k))."""
    a, d1, d2, b = theta.shape
    u, lm, vh = np.linalg.svd(theta.reshape(a * d1, d2 * b), full_matrices=False)
    k = min(chi, a * d1, lm.size)
    return u[:, :k].reshape(a, d1, k), lm[:k], vh[:k].reshape(k, d2, b)


# --------------------------------------------------------------------------------------------------
# The MPS object (MPSClass.py MpsOpenBoundaryClass), cache-free (is_save_op=False semantics)
# --------------------------------------------------------------------------------------------------


class OracleMps:
    """One-site DMRG state with the reference's semantics.  No effective-operator cache: every environment
    is recomputed from the tensors (the reference's is_save_op=False branch, MPSClass.py:352-392), which is
    also the branch whose correlators are correct."""

    def __init__(self, length, d, chi, operators, way='qr', mps=None):
        self.length, self.phys_dim, self.chi = length, d, chi
        self.decomp_way = way
        self.operators = operators
        self.mps = random_open_mps(length, d, chi) if mps is None else [np.array(t) for t in mps]
        self.center = -1
        self.lm = [np.zeros(0) for _ in range(length - 1)]
        self.ent = np.zeros((length - 1, 1))
        self.virtual_dim = np.ones(length + 1, dtype=int) * chi
        self.virtual_dim[0] = self.virtual_dim[-1] = 1
        self.n_matvec = 0
        self.n_solves = 0

    # ---- gauge moves (MPSClass.py:143-198) ----
    def orthogonalize(self, l0, l1):
        if l0 < l1:
            for n in range(l0, l1):
                self.mps[n], mat, self.virtual_dim[n + 1], lm = decompose_l2r(self.mps[n], self.decomp_way)
                if lm.size > 0 and self.center > -1:
                    self.lm[n] = lm.copy()
                self.mps[n + 1] = mode_product(self.mps[n + 1], mat, 0)
        elif l0 > l1:
            for n in range(l0, l1, -1):
                self.mps[n], mat, self.virtual_dim[n], lm = decompose_r2l(self.mps[n], self.decomp_way)
                if lm.size > 0 and self.center > -1:
                    self.lm[n - 1] = lm.copy()
                self.mps[n - 1] = mode_product(self.mps[n - 1], mat, 2)

    def correct_orthogonal_center(self, p):
        if self.center < 0:
            self.orthogonalize(0, p)
            self.orthogonalize(self.length - 1, p)
        elif self.center != p:
            self.orthogonalize(self.center, p)
        self.center = p

    # ---- environments (MPSClass.py:336-441, cache-free branch) ----
    def _chain_l2r(self, l0, l1, v):
        for n in range(l0, l1):
            v = transfer_l2r(self.mps[n], env=v)
        return v

    def _chain_r2l(self, l0, l1, v):
        for n in range(l0, l1, -1):
            v = transfer_r2l(self.mps[n], env=v)
        return v

    def env_one_body(self, p, sn, pos):
        """(vL, vM, vR) for op[sn] at site pos, centre p (environment_s1, MPSClass.py:408-441)."""
        op = self.operators[sn]
        a, d, b = self.mps[p].shape
        if pos > p:
            v = self._chain_r2l(pos - 1, p, transfer_r2l(self.mps[pos], op))
            return np.eye(a), np.eye(d), v
        if pos < p:
            v = self._chain_l2r(pos + 1, p, transfer_l2r(self.mps[pos], op))
            return v, np.eye(d), np.eye(b)
        return np.eye(a), op, np.eye(b)

    def env_two_body(self, p, sn, pos):
        """(vL, vM, vR) for op[sn[0]] at pos[0] and op[sn[1]] at pos[1] (environment_s1_s2,
        MPSClass.py:336-398).  Sites are sorted as the reference does at :345-347."""
        (p0, s0), (p1, s1) = sorted(zip([int(pos[0]), int(pos[1])], [int(sn[0]), int(sn[1])]))
        o0, o1 = self.operators[s0], self.operators[s1]
        a, d, b = self.mps[p].shape
        if p < p0:
            v = transfer_r2l(self.mps[p1], o1)
            v = self._chain_r2l(p1 - 1, p0, v)
            v = transfer_r2l(self.mps[p0], o0, v)
            return np.eye(a), np.eye(d), self._chain_r2l(p0 - 1, p, v)
        if p > p1:
            v = transfer_l2r(self.mps[p0], o0)
            v = self._chain_l2r(p0 + 1, p1, v)
            v = transfer_l2r(self.mps[p1], o1, v)
            return self._chain_l2r(p1 + 1, p, v), np.eye(d), np.eye(b)
        if p == p0:
            v = self._chain_r2l(p1 - 1, p, transfer_r2l(self.mps[p1], o1))
            return np.eye(a), o0, v
        if p == p1:
            v = self._chain_l2r(p0 + 1, p, transfer_l2r(self.mps[p0], o0))
            return v, o1, np.eye(b)
        vl = self._chain_l2r(p0 + 1, p, transfer_l2r(self.mps[p0], o0))
        vr = self._chain_r2l(p1 - 1, p, transfer_r2l(self.mps[p1], o1))
        return vl, np.eye(d), vr

    def grouped_environments(self, p, index1, index2, coeff1, coeff2, tol):
        """The opt_env dictionary of all_environments_optimized + classify_and_update_env
        (MPSClass.py:633-733).  Keys: '1_0_0' (left block), '0_0_1' (right block), '0_s_0' (on-site),
        '1_s_0' / '0_s_1' (left/right block times site operator s), '1_0_1' (list of [c, L, R]).
        Quirk preserved: terms are dropped when |coeff| <= tol with tol = eigs_tol (:645,650,800)."""
        env = {}

        def acc(key, mat):
            env[key] = env[key] + mat if key in env else mat.copy() if hasattr(mat, 'copy') else mat

        for n in range(index1.shape[0]):
            pos, sn = int(index1[n, 0]), int(index1[n, 1])
            c = float(np.ravel(coeff1)[n])
            if abs(c) > tol and np.linalg.norm(self.operators[sn]) > tol:
                vl, vm, vr = self.env_one_body(p, sn, pos)
                if pos < p:
                    acc('1_0_0', vl * c)
                elif pos == p:
                    acc('0_%d_0' % sn, vm * c)
                else:
                    acc('0_0_1', vr * c)
        for n in range(index2.shape[0]):
            c = float(np.ravel(coeff2)[n])
            if abs(c) <= tol:
                continue
            pos = [int(index2[n, 0]), int(index2[n, 1])]
            sn = [int(index2[n, 2]), int(index2[n, 3])]
            assert pos[0] < pos[1], 'the reference path is only correct for site1 < site2 (SURVEY.md section 7)'
            vl, vm, vr = self.env_two_body(p, sn, pos)
            if pos[1] < p:
                acc('1_0_0', vl * c)
            elif pos[0] > p:
                acc('0_0_1', vr * c)
            elif pos[1] == p:
                acc('1_%d_0' % sn[1], vl * c)
            elif pos[0] == p:
                acc('0_%d_1' % sn[0], vr * c)
            else:
                env.setdefault('1_0_1', []).append([c, vl, vr])
        return env

    def apply_handle(self, psi, env, shape, tau):
        """psi -> psi - tau * H_eff psi (update_tensor_eigs_f_handle_optimized, MPSClass.py:755-776)."""
        t = np.asarray(psi).reshape(shape)
        a, d, b = shape
        out = t.copy()

        def left(mat, x):   # E . x on bond 0 (absorb_matrix2tensor(x, E.T, 0))
            return (mat @ x.reshape(a, d * b)).reshape(a, d, b)

        def right(mat, x):  # x . E^T on bond 2
            return (x.reshape(a * d, b) @ mat.T).reshape(a, d, b)

        def mid(op, x):     # op on bond 1
            return np.einsum('st,atb->asb', op, x)

        for key, val in env.items():
            x = key.split('_')
            if key == '1_0_1':
                for c, vl, vr in val:
                    out -= tau * c * right(vr, left(vl, t))
            elif x[0] == '1' and x[1] == '0':
                out -= tau * left(val, t)
            elif x[2] == '1' and x[1] == '0':
                out -= tau * right(val, t)
            elif x[0] == '0' and x[2] == '0':
                out -= tau * mid(val, t)
            elif x[0] == '1':
                out -= tau * left(val, mid(self.operators[int(x[1])], t))
            else:
                out -= tau * right(val, mid(self.operators[int(x[1])], t))
        self.n_matvec += 1
        return out.reshape(-1)

    def dense_effective_hamiltonian(self, p, index1, index2, coeff1, coeff2, tol=1e-12):
        """H_eff = sum c kron(kron(vL, vM), vR)  (effective_hamiltonian_dmrg, MPSClass.py:532-580)."""
        a, d, b = self.mps[p].shape
        h = np.zeros((a * d * b, a * d * b))
        for n in range(index1.shape[0]):
            if abs(coeff1[n]) > tol and np.linalg.norm(self.operators[int(index1[n, 1])]) > tol:
                vl, vm, vr = self.env_one_body(p, int(index1[n, 1]), int(index1[n, 0]))
                h += float(np.ravel(coeff1)[n]) * np.kron(np.kron(vl, vm), vr)
        for n in range(index2.shape[0]):
            if abs(coeff2[n]) > tol:
                vl, vm, vr = self.env_two_body(p, index2[n, 2:4], index2[n, :2])
                h += float(np.ravel(coeff2)[n]) * np.kron(np.kron(vl, vm), vr)
        return h

    # ---- local update (update_tensor_eigs, MPSClass.py:778-809) ----
    def update_tensor_eigs(self, p, index1, index2, coeff1, coeff2, tau, is_real, tol):
        self.correct_orthogonal_center(p)
        shape = self.mps[p].shape
        env = self.grouped_environments(p, index1, index2, coeff1, coeff2, tol)
        dim = int(np.prod(shape))
        h = LinearOperator((dim, dim), matvec=lambda v: self.apply_handle(v, env, shape, tau), dtype=float)
        vec = eigsh(h, k=1, which='LM', v0=self.mps[p].reshape(-1), tol=tol)[1].reshape(shape)
        self.mps[p] = vec.real if is_real else vec
        self.n_solves += 1

    # ---- observables (MPSClass.py:857-952), cache-free branch ----
    def observe_one_body(self, sn, pos):
        op = self.operators[sn]
        if pos > self.center:
            v = self._chain_r2l(pos - 1, self.center - 1, transfer_r2l(self.mps[pos], op))
        else:
            v = self._chain_l2r(pos + 1, self.center + 1, transfer_l2r(self.mps[pos], op))
        return np.trace(v)

    def observe_two_body(self, sn, pos):
        (p0, s0), (p1, s1) = sorted(zip([int(pos[0]), int(pos[1])], [int(sn[0]), int(sn[1])]))
        v = self._chain_l2r(self.center, p0, None) if self.center < p0 else None
        v = transfer_l2r(self.mps[p0], self.operators[s0], v)
        v = self._chain_l2r(p0 + 1, p1, v)
        v = transfer_l2r(self.mps[p1], self.operators[s1], v)
        if p1 < self.center:
            v = self._chain_l2r(p1 + 1, self.center + 1, v)
        return np.trace(v)

    def observe_magnetization(self, sn):
        return np.array([[self.observe_one_body(sn, i)] for i in range(self.length)]).real

    def observe_bond_energy(self, index2, coeff2):
        return np.array([[float(np.ravel(coeff2)[n]) * self.observe_two_body(index2[n, 2:], index2[n, :2])]
                         for n in range(index2.shape[0])]).real

    def observe_correlators_from_middle(self, op1, op2, ob_len=None):
        """walks outwards from round(L/2) alternating left / right (MPSClass.py:936-952)."""
        ob_len = self.length if ob_len is None else ob_len
        mid = round(self.length / 2)
        pos1, pos2, k, out = mid, mid + 1, 0, []
        while pos1 > -0.1 and pos2 < self.length and ob_len > -0.1:
            out.append(self.observe_two_body([op1, op2], [pos1, pos2]))
            if k % 2 == 0:
                pos1 -= 1
            else:
                pos2 += 1
            k += 1
            ob_len -= 1
        return np.array(out).real

    # ---- spectrum (MPSClass.py:812-839) ----
    def calculate_entanglement_spectrum(self):
        way, center = self.decomp_way, self.center
        self.decomp_way = 'svd'
        empty = [n for n in range(self.length - 1) if self.lm[n].size == 0]
        p0, p1 = (min(empty), max(empty)) if empty else (self.length - 1, 0)
        self.correct_orthogonal_center(p0)
        self.correct_orthogonal_center(p1 + 1)
        self.correct_orthogonal_center(center)
        self.decomp_way = way

    def calculate_entanglement_entropy(self):
        for i in range(self.length - 1):
            self.ent[i] = -1 if self.lm[i].size == 0 else entanglement_entropy(self.lm[i])

    def norm(self):
        return float(np.linalg.norm(self.mps[self.center]))


def truncate_mps(mps, chi):
    """library/MPSClass.py:186-247 with is_trun=True (as driven by truncate_virtual_bonds :909-923, way != 'simple'):
    centre at site 0, then an SVD sweep 0 -> L-1 that keeps the chi largest singular triplets of every bond and normalises
    the tensor absorbing the remainder.  Returns (tensors with the centre at L-1, kept singular values per bond)."""
    A = OracleMps(len(mps), mps[0].shape[1], max(t.shape[2] for t in mps), None, way='qr', mps=mps)
    A.correct_orthogonal_center(0)
    out, lms = [np.array(t) for t in A.mps], []
    for n in range(len(out) - 1):
        a, d, b = out[n].shape
        u, lm, vh = np.linalg.svd(out[n].reshape(a * d, b), full_matrices=False)
        k = min(chi, lm.size)
        out[n] = u[:, :k].reshape(a, d, k)
        out[n + 1] = np.einsum('kb,bsc->ksc', lm[:k, None] * vh[:k], out[n + 1])
        out[n + 1] /= np.linalg.norm(out[n + 1])
        lms.append(lm[:k])
    return out, lms


def mps_overlap(x, y):
    """<x|y> of two open-boundary MPS (lists of (a,d,b) tensors)"""
    e = np.ones((1, 1))
    for tx, ty in zip(x, y):
        e = np.einsum('ab,asc,bsd->cd', e, tx.conj(), ty)
    return e[0, 0]


def bond_energies(eb_full, positions, index2):
    """sum the per-term energies onto their bond (get_bond_energies, DMRG_anyH.py:250-258)."""
    eb = np.zeros((positions.shape[0], 1))
    for i in range(eb_full.size):
        hit = (index2[i, 0] == positions[:, 0]) & (index2[i, 1] == positions[:, 1])
        eb[hit] += eb_full[i]
    return eb


def dmrg_finite_size(para, seed=None, max_updates=None):
    """One-site finite DMRG driver, the control flow of DMRG_anyH.py:18-101.
    Returns (ob, A, info).  `max_updates` bounds the number of local updates (bench sampling only)."""
    import time
    t0 = time.time()
    if seed is not None:
        np.random.seed(seed)
    L = para['l']
    A = OracleMps(L, para['d'], para['chi'], para['op'], way='qr')
    A.correct_orthogonal_center(para['ob_position'])
    args = (para['index1'], para['index2'], para['coeff1'], para['coeff2'], para['tau'], para['is_real'],
            para['eigs_tol'])
    ob, info, e0 = {}, {'convergence': 1, 'n_updates': 0}, 0.0
    stop = False
    for t in range(para['sweep_time']):
        if_ob = ((t + 1) % para['dt_ob'] == 0) or t == para['sweep_time'] - 1
        order = list(range(para['ob_position'] + 1, L)) + list(range(L - 2, -1, -1)) + \
            list(range(1, para['ob_position']))
        for n in order:
            A.update_tensor_eigs(n, *args)
            info['n_updates'] += 1
            if max_updates is not None and info['n_updates'] >= max_updates:
                stop = True
                break
        if stop:
            break
        if if_ob:
            ob['eb_full'] = A.observe_bond_energy(para['index2'], para['coeff2'])
            ob['mx'] = A.observe_magnetization(1)
            ob['mz'] = A.observe_magnetization(3)
            ob['e_per_site'] = (ob['eb_full'].sum(axis=0) - para['hx'] * ob['mx'].sum(axis=0) -
                                para['hz'] * ob['mz'].sum(axis=0)) / L
            info['convergence'] = abs(float(ob['e_per_site'][0]) - e0)
            if info['convergence'] < para['break_tol']:
                break
            e0 = float(ob['e_per_site'][0])
    info['t_sweeps'] = time.time() - t0
    if not stop:
        ob['eb'] = bond_energies(ob['eb_full'], para['positions_h2'], para['index2'])
        A.calculate_entanglement_spectrum()
        A.calculate_entanglement_entropy()
        ob['corr_x'] = A.observe_correlators_from_middle(1, 1)
        ob['corr_z'] = A.observe_correlators_from_middle(3, 3)
    info['t_cost'] = time.time() - t0
    info['n_matvec'], info['n_solves'] = A.n_matvec, A.n_solves
    return ob, A, info


# --------------------------------------------------------------------------------------------------
# Independent known-answer check: dense exact diagonalisation of the same term list
# --------------------------------------------------------------------------------------------------


def dense_hamiltonian(para):
    """Full 2^L Hamiltonian sum_n c1 op(site) + sum_n c2 op1(site1) op2(site2) built from the para dict
    (same meaning as index1/index2/coeff1/coeff2 in Parameters.py:36-52).  Small L only."""
    L, d, ops = para['l'], para['d'], para['op']
    dim = d ** L

    def site_op(o, i):
        return np.kron(np.kron(np.eye(d ** i), o), np.eye(d ** (L - i - 1)))

    h = np.zeros((dim, dim))
    for n in range(para['index1'].shape[0]):
        o = np.real(ops[int(para['index1'][n, 1])])
        if np.linalg.norm(o) > 0:
            h += float(np.ravel(para['coeff1'])[n]) * site_op(o, int(para['index1'][n, 0]))
    for n in range(para['index2'].shape[0]):
        i, j, s1, s2 = [int(x) for x in para['index2'][n]]
        h += float(np.ravel(para['coeff2'])[n]) * site_op(np.real(ops[s1]), i) @ site_op(np.real(ops[s2]), j)
    return h


# --------------------------------------------------------------------------------------------------
# Two-site update (north_star kernels 1 + 3).  The reference's finite driver is one-site only; its two-site
# machinery (library/MPSClass.py:1676-1707: theta contraction, term-summed matvec, SVD truncation) is used by iDMRG.
# The oracle states the finite-size two-site algorithm through the DENSE Hamiltonian, so it shares no code path
# with the product's grouped environments: H_eff = P^T H P with P the isometry of the two-site window.
# --------------------------------------------------------------------------------------------------


def two_site_isometry(mps, p):
    """P[(s_0 .. s_{L-1}), (a, s_p, s_{p+1}, b)] for an MPS whose sites < p are left- and sites > p+1 right-orthogonal."""
    L, d = len(mps), mps[0].shape[1]
    left = np.ones((1, 1))                      # (config of sites < p, a)
    for n in range(p):
        t = mps[n]
        left = np.einsum('ca,asb->csb', left, t).reshape(-1, t.shape[2])
    right = np.ones((1, 1))                     # (b, config of sites > p+1)
    for n in range(L - 1, p + 1, -1):
        t = mps[n]
        right = np.einsum('asb,bc->asc', t, right).reshape(t.shape[0], -1)
    a, b = left.shape[1], right.shape[0]
    eye = np.eye(d * d).reshape(d, d, d, d)     # (s_p s_{p+1}; s s')
    iso = np.einsum('ca,xyst,bf->cxyfastb', left, eye, right)
    return iso.reshape(d ** L, a * d * d * b)


def dense_two_site_effective_hamiltonian(mps, p, para):
    iso = two_site_isometry(mps, p)
    return iso.T @ dense_hamiltonian(para) @ iso


def dmrg_two_site_dense(para, chi, n_sweeps, seed=0, chi_init=2):
    """finite two-site DMRG with dense local solves (small L): returns (energy, mps, lm list)."""
    rng = np.random.RandomState(seed)
    L, d = para['l'], para['d']
    dims = [1] + [chi_init] * (L - 1) + [1]
    mps = [rng.randn(dims[n], d, dims[n + 1]) for n in range(L)]
    for n in range(L - 1, 0, -1):               # right-orthogonalise: centre at 0
        mps[n], r, _, _ = decompose_r2l(mps[n], 'qr')
        mps[n - 1] = mode_product(mps[n - 1], r, 2)
    h = dense_hamiltonian(para)
    lms, energy = [None] * (L - 1), None

    def solve(p, to_right):
        nonlocal energy
        iso = two_site_isometry(mps, p)
        w, v = np.linalg.eigh(iso.T @ h @ iso)
        energy = w[0]
        a, b = mps[p].shape[0], mps[p + 1].shape[2]
        u, lm, vh = svd_truncate_two_site(v[:, 0].reshape(a, d, d, b), chi)
        lm = lm / np.linalg.norm(lm)
        lms[p] = lm
        if to_right:
            mps[p], mps[p + 1] = u, lm[:, None, None] * vh
        else:
            mps[p], mps[p + 1] = u * lm[None, None, :], vh

    for _ in range(n_sweeps):
        for p in range(L - 1):
            solve(p, True)
        for p in range(L - 2, -1, -1):
            solve(p, False)
    return energy, mps, lms
