"""Oracle-backed stand-in for tnalg_b200.ops.CudaBackend -- TEST INFRASTRUCTURE ONLY.

Implements the same interface with numpy (via oracle/dmrg_oracle.py primitives) on CPU torch tensors so that the
host logic of the product (term algebra, environment bookkeeping, sweep driver, observables, pickling) can be
exercised by `-m "not gpu"` tests in a container without a GPU.  The product never imports this module and has no
hook for it: tests replace the process-wide backend object with `install()` below.

`lanczos` is a numpy transcription of the algorithm in tnalg_b200/csrc/lanczos.cu (thick restart, CGS2, ARPACK
criterion on the shifted operator) so that its convergence behaviour is validated against the reference's results.
"""
import numpy as np
import torch
from scipy.linalg import eigh_tridiagonal

from oracle import dmrg_oracle as orc


def install(be):
    """make `be` the process-wide backend returned by tnalg_b200.ops.backend() (test-side injection; the product has no
    backend switch)"""
    from tnalg_b200 import ops
    ops._backend = be


class GlooComm:
    """stand-in for tnalg_b200.ops.Comm (the in-library NCCL communicator) on torch.distributed/gloo"""

    def __init__(self, dist):
        self.dist, self.rank, self.world = dist, dist.get_rank(), dist.get_world_size()

    def allreduce(self, t):
        self.dist.all_reduce(t)
        return t

    def broadcast(self, t, src=0):
        self.dist.broadcast(t, src=src)
        return t

    def broadcast_many(self, tensors, roots):
        for t, r in zip(tensors, roots):
            self.dist.broadcast(t, src=r)

    def allgather_inplace(self, buf):
        per = buf.shape[0] // self.world
        parts = [torch.empty_like(buf[:per]) for _ in range(self.world)]
        self.dist.all_gather(parts, buf[self.rank * per:(self.rank + 1) * per].clone())
        for r, part in enumerate(parts):
            buf[r * per:(r + 1) * per] = part
        return buf


def shard_groups(g, rank, world):
    """round-robin ownership of the links in the order HL, LS.., HR, RS.., X.. (tn_effh_plan_create); the on-site part
    belongs to rank 0"""
    if world <= 1:
        return g
    out = {k: (list(v) if isinstance(v, list) else v) for k, v in g.items()}
    idx = 0

    def own():
        nonlocal idx
        mine = idx % world == rank
        idx += 1
        return mine
    if not own():
        out['HL'] = None
    keep = [own() for _ in g['LS']]
    out['LS'] = [x for x, k in zip(g['LS'], keep) if k]
    out['ls_ops'] = [x for x, k in zip(g['ls_ops'], keep) if k]
    if not own():
        out['HR'] = None
    keep = [own() for _ in g['RS']]
    out['RS'] = [x for x, k in zip(g['RS'], keep) if k]
    out['rs_ops'] = [x for x, k in zip(g['rs_ops'], keep) if k]
    keep = [own() for _ in g['XL']]
    for key in ('XL', 'XR', 'x_coeff'):
        out[key] = [x for x, k in zip(g[key], keep) if k]
    if rank != 0:
        out['M'] = None
    return out


class CpuPlan:
    def __init__(self, shape, g, rank=0, world=1):
        self.shape = shape
        self.rank, self.world = rank, world
        a, d, b = shape
        full = g
        g = shard_groups(g, rank, world)
        self.g = g
        self._full = full
        kl = (1 if full['HL'] is not None else 0) + len(full['LS'])
        kr = (1 if full['HR'] is not None else 0) + len(full['RS'])
        nx = len(full['XL'])
        self.flops_algorithmic = 2.0 * a * d * b * (a * (kl + nx) + b * (kr + nx))
        self.flops_executed = 2.0 * a * d * b * (a * ((1 if g['HL'] is not None else 0) + len(g['LS']) + len(g['XL'])) +
                                                 b * ((1 if g['HR'] is not None else 0) + len(g['RS']) + len(g['XL'])))
        self._handle = self
        self.uses_tma = 0

    def apply(self, x):
        """H_eff x as GEMMs (the formulation of absorb_matrix2tensor, TensorBasicModule.py:387-424)"""
        g = self.g
        a, d, b = self.shape
        t = np.asarray(x).reshape(self.shape)

        def left(mat, y):
            return (mat @ y.reshape(a, d * b)).reshape(a, d, b)

        def right(mat, y):
            return (y.reshape(a * d, b) @ mat.T).reshape(a, d, b)

        def mid(op, y):
            return np.einsum('st,atb->asb', op, y)

        out = np.zeros_like(t)
        if g['HL'] is not None:
            out += left(g['HL'], t)
        if g['HR'] is not None:
            out += right(g['HR'], t)
        if g['M'] is not None:
            out += mid(g['M'], t)
        for E, op in zip(g['LS'], g['ls_ops']):
            out += left(E, mid(op, t))
        for E, op in zip(g['RS'], g['rs_ops']):
            out += right(E, mid(op, t))
        for c, l, r in zip(g['x_coeff'], g['XL'], g['XR']):
            out += c * right(r, left(l, t))
        return out.reshape(-1)

    def matvec(self, psi, c_id=0.0, c_h=1.0, out=None):
        x = psi.numpy().reshape(-1)
        y = (c_id * x if self.rank == 0 else 0.0) + c_h * self.apply(x)
        res = torch.from_numpy(y.reshape(psi.shape))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def destroy(self):
        pass


class CpuBackend:
    name = 'cpu-oracle'
    device = torch.device('cpu')
    sm_count = 0

    def __init__(self):
        self.n_env_updates = 0
        self.n_matvec = 0
        self.last_svd_sweeps = 0

    def from_numpy(self, x):
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))

    def to_numpy(self, t):
        return t.detach().numpy().copy()

    def to_numpy_many(self, tensors):
        return [self.to_numpy(t) for t in tensors]

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64)

    def launch_count(self):
        return 0

    def comm(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return GlooComm(dist)
        return None

    def env_update(self, direction, T, outputs, outs=None):
        self.n_env_updates += 1
        Tn = T.numpy()
        res = []
        fn = orc.transfer_l2r if direction == 0 else orc.transfer_r2l
        for links in outputs:
            acc = None
            for E, op in links:
                v = fn(Tn, None if op is None else np.asarray(op), None if E is None else E.numpy())
                acc = v if acc is None else acc + v
            res.append(torch.from_numpy(np.ascontiguousarray(acc)))
        if outs is not None:
            for o, r in zip(outs, res):
                o.copy_(r)
            return outs
        return res

    def lincomb(self, xs, coeffs):
        return sum(float(c) * x for c, x in zip(coeffs, xs))

    def site_op(self, T, op):
        return torch.from_numpy(np.einsum('st,atb->asb', np.asarray(op), T.numpy()))

    def dot(self, x, y):
        return torch.tensor([float((x * y).sum())], dtype=torch.float64)

    def trace(self, E):
        return torch.tensor([float(torch.trace(E))], dtype=torch.float64)

    def scalars_to_host(self, slots):
        return torch.cat(slots).numpy() if slots else np.zeros(0)

    def mode_product(self, T, mat, bond):
        return torch.from_numpy(np.ascontiguousarray(orc.mode_product(T.numpy(), mat.numpy(), bond)))

    def effh_plan(self, shape, HL=None, HR=None, M=None, LS=(), ls_ops=(), RS=(), rs_ops=(), XL=(), XR=(), x_coeff=(),
                  rank=0, world=1, rows=None):
        assert rows is None, 'the stand-in shards by terms only'
        n = lambda t: None if t is None else t.numpy()  # noqa: E731
        g = {'HL': n(HL), 'HR': n(HR), 'M': None if M is None else np.asarray(M), 'LS': [n(t) for t in LS],
             'ls_ops': [np.asarray(o) for o in ls_ops], 'RS': [n(t) for t in RS], 'rs_ops': [np.asarray(o) for o in rs_ops],
             'XL': [n(t) for t in XL], 'XR': [n(t) for t in XR], 'x_coeff': list(x_coeff)}
        return CpuPlan(tuple(shape), g, rank, world)

    def lanczos(self, plan, tau, v0, tol, ncv=20, max_restarts=1000, comm=None, allreduce=None):
        """numpy transcription of tn_lanczos_lm1 (tnalg_b200/csrc/lanczos.cu)."""
        x0 = v0.numpy().reshape(-1)
        n = x0.size
        m = int(min(max(ncv, 2), n, 64))
        V = np.zeros((m + 1, n))
        V[0] = x0 / np.linalg.norm(x0)
        alpha, beta = np.zeros(m), np.zeros(m)
        j0, n_mv = 0, 0
        eps23 = 3.666852862501036e-11
        y, lam, resid, ok = V[0], 0.0, 0.0, False
        for cycle in range(max_restarts):
            m_eff = m
            for j in range(j0, m):
                w = np.ascontiguousarray(plan.apply(V[j]))
                n_mv += 1
                if comm is not None:
                    w = comm.allreduce(torch.from_numpy(w)).numpy()
                elif allreduce is not None:
                    assert allreduce(w.ctypes.data, w.size, None, None) == 0
                h = V[:j + 1] @ w
                w = w - V[:j + 1].T @ h
                h2 = V[:j + 1] @ w
                w = w - V[:j + 1].T @ h2
                alpha[j] = h[j] + h2[j]
                bj = np.linalg.norm(w)
                beta[j] = bj
                scale = max(abs(alpha[j]), abs(beta[j - 1]) if j > 0 else 0.0, 1e-300)
                if bj <= 1e-14 * scale:
                    beta[j] = 0.0
                    m_eff = j + 1
                    break
                V[j + 1] = w / bj
            breakdown = m_eff < m or beta[m_eff - 1] == 0.0
            d, z = eigh_tridiagonal(alpha[:m_eff], beta[:m_eff - 1]) if m_eff > 1 else (alpha[:1].copy(), np.ones((1, 1)))
            best = int(np.argmax(np.abs(1.0 - tau * d)))
            u = z[:, best]
            s = (0.0 if breakdown else beta[m_eff - 1]) * u[m_eff - 1]
            theta, lam, resid = d[best], 1.0 - tau * d[best], abs(s)
            y = u @ V[:m_eff]
            if breakdown or abs(tau) * abs(s) <= tol * max(eps23, abs(lam)) or m >= n:
                ok = True
                break
            r_hat = V[m].copy()
            V[0], V[1] = y, r_hat
            alpha[0], beta[0] = theta, s
            j0 = 1
        self.n_matvec += n_mv
        y = y / np.linalg.norm(y)
        return lam, torch.from_numpy(y.copy()), n_mv, resid, ok

    def svd(self, A, k_keep=None):
        u, s, vt = np.linalg.svd(A.numpy(), full_matrices=False)
        k = s.size if k_keep is None else min(k_keep, s.size)
        c = np.ascontiguousarray
        return torch.from_numpy(c(u[:, :k])), torch.from_numpy(c(s[:k])), torch.from_numpy(c(vt[:k]))

    def qr(self, A):
        q, r = np.linalg.qr(A.numpy())
        return torch.from_numpy(np.ascontiguousarray(q)), torch.from_numpy(np.ascontiguousarray(r))

    def qr_tensor(self, T, left2right):
        a, d, b = T.shape
        if left2right:
            q, r = self.qr(T.reshape(a * d, b))
            return q.reshape(a, d, -1).contiguous(), r
        q, r = self.qr(T.reshape(a, d * b).t().contiguous())
        return q.t().contiguous().reshape(-1, d, b), r

    def eigh(self, A):
        w, v = np.linalg.eigh(A.numpy())
        return torch.from_numpy(w), torch.from_numpy(np.ascontiguousarray(v))

    def lanczos_generic(self, matvec, n, tau, v0, tol, ncv=20, max_restarts=1000, locked=None):
        """dense stand-in: build the operator column by column, deflate, diagonalise"""
        H = np.zeros((n, n))
        for j in range(n):
            e = torch.zeros(n, dtype=torch.float64)
            e[j] = 1.0
            y = torch.zeros(n, dtype=torch.float64)
            matvec(e, y)
            H[:, j] = y.numpy()
        H = 0.5 * (H + H.T)
        basis = np.eye(n)
        if locked is not None and locked.shape[0]:
            q, _ = np.linalg.qr(locked.numpy().T, mode='complete')
            basis = q[:, locked.shape[0]:]                 # orthogonal complement of the locked vectors
        w, v = np.linalg.eigh(basis.T @ H @ basis)
        best = int(np.argmax(np.abs(1.0 - tau * w)))
        return 1.0 - tau * w[best], torch.from_numpy(np.ascontiguousarray(basis @ v[:, best])), n, 0.0, True

    def scale_diag_rows(self, S, Vt):
        return S[:, None] * Vt

    def norm(self, x):
        return float(torch.linalg.vector_norm(x))
