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
