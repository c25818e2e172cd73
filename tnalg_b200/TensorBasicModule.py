"""Tensor primitives of the hot path with the reference's names and argument meaning
(TensorBasicModule.py:181-186,314-424,530-619,786-801), executed by the CUDA library.
Inputs may be numpy arrays or CUDA tensors; numpy in -> numpy out."""
import numpy as np

from . import ops as _ops


def _dev(x, be):
    if x is None:
        return None, False
    if hasattr(x, 'data_ptr'):
        return x, True
    x = np.asarray(x)
    if x.size == 0:
        return None, False
    return be.from_numpy(np.real(x)), False


def _ret(t, be, as_tensor):
    return t if as_tensor else be.to_numpy(t)


def random_open_mps(l, d, chi):
    mps = [None] * l
    mps[0] = np.random.randn(1, d, chi)
    mps[l - 1] = np.random.randn(chi, d, 1)
    for n in range(1, l - 1):
        mps[n] = np.random.randn(chi, d, chi)
    return mps


def absorb_matrix2tensor(tensor, mat, bond):
    """out[.., j, ..] = sum_i tensor[.., i, ..] mat[i, j] for a rank-3 tensor (TensorBasicModule.py:387-424)."""
    be = _ops.backend()
    t, tt = _dev(tensor, be)
    m, _ = _dev(mat, be)
    return _ret(be.mode_product(t, m, bond), be, tt)


def _op_host(op):
    if op is None:
        return None
    if hasattr(op, 'data_ptr'):
        op = op.cpu().numpy()
    op = np.asarray(op)
    return None if op.size == 0 else np.real(op).astype(float)


def bound_vec_operator_left2right(tensor, op=np.zeros(0), v=np.zeros(0), normalize=False, symme=False):
    """E'[b,b'] = sum conj(T[a,s,b]) v[a,a'] op[s,s'] T[a',s',b'] (TensorBasicModule.py:530-573)."""
    be = _ops.backend()
    t, tt = _dev(tensor, be)
    e, _ = _dev(v, be)
    out = be.env_update(0, t, [[(e, _op_host(op))]])[0]
    return _post(out, be, tt, normalize, symme)


def bound_vec_operator_right2left(tensor, op=np.zeros(0), v=np.zeros(0), normalize=False, symme=False):
    """E'[a,a'] = sum conj(T[a,s,b]) v[b,b'] op[s,s'] T[a',s',b'] (TensorBasicModule.py:576-619)."""
    be = _ops.backend()
    t, tt = _dev(tensor, be)
    e, _ = _dev(v, be)
    out = be.env_update(1, t, [[(e, _op_host(op))]])[0]
    return _post(out, be, tt, normalize, symme)


def _post(out, be, as_tensor, normalize, symme):
    if normalize:
        out = out / be.norm(out)
    if symme:
        out = (out + out.t()) / 2
    return _ret(out, be, as_tensor)


def left2right_decompose_tensor(tensor, way='qr', is_full=False):
    """(a,d,b) -> Q (a,d,k), v = R^T (b,k), k, lm (TensorBasicModule.py:314-348)."""
    be = _ops.backend()
    t, tt = _dev(tensor, be)
    a, d, b = t.shape
    mat = t.reshape(a * d, b)
    k = min(a * d, b)
    if way == 1 or way == 'svd':
        U, S, Vt = be.svd(mat)
        Q, R, lm = U, be.scale_diag_rows(S, Vt), be.to_numpy(S)
    else:
        Q, R = be.qr(mat)
        lm = np.zeros(0)
    return _ret(Q.contiguous().reshape(a, d, k), be, tt), _ret(R.t().contiguous(), be, tt), k, lm


def right2left_decompose_tensor(tensor, way='qr', is_full=False):
    """(a,d,b) -> Q (k,d,b), v = R^T (a,k), k, lm (TensorBasicModule.py:351-384)."""
    be = _ops.backend()
    t, tt = _dev(tensor, be)
    a, d, b = t.shape
    mat = t.reshape(a, d * b).t().contiguous()
    k = min(a, d * b)
    if way == 1 or way == 'svd':
        U, S, Vt = be.svd(mat)
        Q, R, lm = U, be.scale_diag_rows(S, Vt), be.to_numpy(S)
    else:
        Q, R = be.qr(mat)
        lm = np.zeros(0)
    return _ret(Q.t().contiguous().reshape(k, d, b), be, tt), _ret(R.t().contiguous(), be, tt), k, lm


def svd_truncate_two_site(theta, chi):
    """two-site wavefunction (a,d,d,b) -> U (a,d,k), lm (k,), Vh (k,d,b) with k = min(chi, a*d, d*b): the SVD
    truncation of library/MPSClass.py:1676-1686 on the Jacobi kernel."""
    be = _ops.backend()
    t, tt = _dev(theta, be)
    a, d1, d2, b = t.shape
    k = min(chi, a * d1, d2 * b)
    U, S, Vt = be.svd(t.reshape(a * d1, d2 * b), k_keep=k)
    return _ret(U.reshape(a, d1, k), be, tt), be.to_numpy(S), _ret(Vt.reshape(k, d2, b), be, tt)


def entanglement_entropy(lm, tol=1e-20):
    """-2 sum lm^2 ln lm over lm > tol (TensorBasicModule.py:786-801); host arithmetic on chi numbers."""
    lm = np.sort(np.asarray(lm, dtype=float).reshape(-1))[::-1]
    lm = lm[lm > tol]
    return float(-2 * np.dot(lm ** 2, np.log(lm)))


def sort_vectors(mat, order, which='row'):
    mat = np.asarray(mat)
    return mat[order, :] if which == 'row' else mat[:, order]


# ---- a6: several mode products at once (TensorBasicModule.py:427-528) ----
def absorb_matrices2tensor_full_fast(tensor, mats):
    """out[i', j', k'] = sum tensor[i, j, k] mats[0][i, i'] mats[1][j, j'] mats[2][k, k'] for a rank-3 tensor: one mode
    product per bond (chain GEMM / site-operator kernel), first index of every matrix contracted."""
    out = tensor
    for bond, m in enumerate(mats):
        out = absorb_matrix2tensor(out, m, bond)
    return out


def absorb_matrices2tensor(tensor, mats, bonds=np.zeros(0), mat_bond=-1):
    """mode products on the listed bonds (default: all bonds in order); mat_bond[i] == 1 contracts the second index of
    mats[i] instead of the first."""
    mats = list(mats)
    bonds = np.asarray(bonds, dtype=int).reshape(-1)
    if bonds.size == 0:
        bonds = np.arange(len(mats))
    if not np.isscalar(mat_bond):
        flags = np.asarray(mat_bond).reshape(-1)
        mats = [m.T if flags[i] == 1 else m for i, m in enumerate(mats)]
    out = tensor
    for m, b in zip(mats, bonds):
        out = absorb_matrix2tensor(out, m, int(b))
    return out


# ---- transfers that keep one physical bond open (TensorBasicModule.py:622-649; used by the two-body density matrix) ----
def _unit_ops(d):
    ops = []
    for s in range(d):
        for sp in range(d):
            u = np.zeros((d, d))
            u[s, sp] = 1.0
            ops.append(u)
    return ops


def _bound_vec_with_phys(direction, tensor, v, normalize):
    be = _ops.backend()
    t, tt = _dev(tensor, be)
    a, d, b = t.shape
    e_out = b if direction == 0 else a
    import torch
    if v is None or (not hasattr(v, 'data_ptr') and np.asarray(v).size == 0):
        outputs = [[(None, u)] for u in _unit_ops(d)]                # out[s, s'] = sum conj(T[., s, .]) T[., s', .]
    elif (v.dim() if hasattr(v, 'data_ptr') else np.asarray(v).ndim) == 2:
        e, _ = _dev(v, be)
        outputs = [[(e, u)] for u in _unit_ops(d)]                   # ... with the environment v in between
    else:
        e, _ = _dev(np.asarray(v).reshape(d * d, *np.asarray(v).shape[2:]) if not hasattr(v, 'data_ptr') else v.reshape(d * d, *v.shape[2:]), be)
        outputs = [[(e[k].contiguous(), None)] for k in range(d * d)]  # the open bond is already inside v[s, s']
    res = be.env_update(direction, t, outputs)
    out = torch.stack([r for r in res]).reshape(d, d, e_out, e_out)
    if normalize:
        out = out / be.norm(out)
    return _ret(out, be, tt)


def bound_vec_with_phys_left2right(tensor, v=np.zeros(0), normalize=False):
    """out[s, s', b, b'] = sum conj(T[a, s, b]) v[a, a'] T[a', s', b'] (v omitted: identity; 4-index v[s, s', a, a']: plain
    transfer of every (s, s') block).  One batched tn_env_update call with the d*d matrix units as site operators."""
    return _bound_vec_with_phys(0, tensor, v, normalize)


def bound_vec_with_phys_right2left(tensor, v=np.zeros(0), normalize=False):
    """mirror image: out[s, s', a, a'] = sum conj(T[a, s, b]) v[b, b'] T[a', s', b']"""
    return _bound_vec_with_phys(1, tensor, v, normalize)


# ---- small host-side helpers with the reference's names (not on the hot path) ----
def transfer_matrix_mps(tensor):
    """(a*a, b*b) transfer matrix sum_s conj(T[a,s,b]) T[a',s,b'] (TensorBasicModule.py:652-674)"""
    t = tensor.cpu().numpy() if hasattr(tensor, 'data_ptr') else np.asarray(tensor)
    a, _, b = t.shape
    return np.einsum('asb,csd->acbd', t.conj(), t).reshape(a * a, b * b)


def normalize_tensor(tensor, if_flatten=False, is_enforce=False):
    """(tensor / norm, norm) (TensorBasicModule.py:755-785); a norm below 1e-30 leaves the tensor untouched unless is_enforce"""
    t = tensor.cpu().numpy() if hasattr(tensor, 'data_ptr') else np.asarray(tensor)
    v = t.reshape(-1)
    norm = np.linalg.norm(v)
    if norm < 1e-30 and not is_enforce:
        return (v if if_flatten else t), norm
    return (v / norm if if_flatten else t / norm), norm


def is_identity(mat, tol=1e-15, sample_t=10):
    """True when mat = c * identity with c != 0 (TensorBasicModule.py:826-860; checked exactly here, not by sampling)"""
    m = np.asarray(mat)
    if m.ndim != 2 or m.shape[0] != m.shape[1] or abs(m[0, 0]) < tol:
        return False
    return bool(np.abs(m - m[0, 0] * np.eye(m.shape[0])).max() <= max(tol, 1e-15) * max(1.0, abs(m[0, 0])))


def check_orthogonality(tensor, ind0, tol=1e-20):
    """is the tensor an isometry from the remaining indexes onto the indexes ind0? (TensorBasicModule.py:876-900)"""
    t = tensor.cpu().numpy() if hasattr(tensor, 'data_ptr') else np.asarray(tensor)
    ind0 = list(ind0)
    ind1 = [n for n in range(t.ndim) if n not in ind0]
    m = t.transpose(ind0 + ind1).reshape(int(np.prod([t.shape[n] for n in ind0])), -1)
    return is_identity(m.conj().dot(m.T), tol=max(tol, 1e-15))


def ones_open_mps(l, d, chi):
    mps = [np.ones((chi, d, chi)) for _ in range(l)]
    mps[0] = np.ones((1, d, chi))
    mps[l - 1] = np.ones((chi, d, 1))
    return mps
