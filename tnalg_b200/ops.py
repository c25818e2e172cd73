"""Host-side operator layer: torch CUDA tensors in, C-ABI calls out.

`backend()` returns the process-wide CudaBackend.  Every method maps 1:1 onto an entry point of
include/tnalg_b200.h; torch is used for device memory, streams and (multi-GPU) torch.distributed only.
The one library call kept in round 1 is the QR factorisation of gauge moves (torch.linalg.qr -> cuSOLVER geqrf),
see DESIGN.md; everything else runs in libtnalg_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_backend = None


def backend():
    global _backend
    if _backend is None:
        _backend = CudaBackend()
    return _backend


def set_backend(b):
    """Dependency-injection hook used by the CPU unit tests of the host logic (tests/cpu_backend.py)."""
    global _backend
    _backend = b


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _op_array(ops, d):
    """list of (d,d) numpy arrays -> flat ctypes double array"""
    flat = np.zeros(len(ops) * d * d)
    for i, o in enumerate(ops):
        flat[i * d * d:(i + 1) * d * d] = np.asarray(o, dtype=float).reshape(-1)
    return (C.c_double * flat.size)(*flat)


class EffHPlan:
    """Handle on a tn_effh_plan (include/tnalg_b200.h).  Owns the plan workspace (tables + crossing scratch)."""

    def __init__(self, be, handle, workspace, shape, keep):
        self._be, self._handle, self._ws, self.shape, self._keep = be, handle, workspace, shape, keep
        alg, ex = C.c_double(), C.c_double()
        L.check(be.lib.tn_effh_plan_flops(handle, C.byref(alg), C.byref(ex)))
        self.flops_algorithmic, self.flops_executed = alg.value, ex.value
        self.uses_tma = int(be.lib.tn_effh_plan_uses_tma(handle))  # bit 0 left stage, bit 1 right stage

    def matvec(self, psi, c_id=0.0, c_h=1.0, out=None):
        """out = c_id*psi + c_h*H_eff psi; (c_id, c_h) = (1, -tau) is the reference handle (MPSClass.py:755-776)."""
        be = self._be
        psi = psi.contiguous()
        out = torch.empty_like(psi) if out is None else out
        L.check(be.lib.tn_effh_matvec(self._handle, _ptr(psi), _ptr(out), float(c_id), float(c_h), be.stream()))
        return out

    def destroy(self):
        if self._handle is not None:
            self._be.lib.tn_effh_plan_destroy(self._handle)
            self._handle = None
            self._ws = None
            self._keep = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def preconditioned_svd(A, k, qr, jacobi, mm_nn, mm_nt):
    """Truncated SVD of a tall A (m >= n) around a one-sided Jacobi kernel, written against callables so that the algebra is
    unit-tested on the CPU (tests/test_host_logic_cpu.py) with the very code the CUDA backend runs:
        P1: columns of A sorted by decreasing norm,          A P1 = Q1 R1
        P2: rows of R1 sorted by decreasing norm,            (P2 R1)^T = Q2 R2
        Jacobi on the columns of R2^T:                       R2^T = Ux S Vx^T
        =>  A = (Q1 P2^T Ux) S (P1 Q2 Vx)^T
    qr(X) -> (Q, R) thin; jacobi(X, k) -> (U (n,k), S (k), Vt (k,n)); mm_nn(X, Y) = X Y; mm_nt(X, Y) = X Y^T.
    Sorting by norm is the cheap stand-in for a column-pivoted QR: on two-site DMRG wavefunctions (strongly graded, with
    SU(2) multiplets) it cuts the Jacobi sweeps from 11-13 to 6-8; on synthetic graded matrices it changes nothing."""
    import torch
    p1 = torch.argsort((A * A).sum(0), descending=True)
    Q1, R1 = qr(A.index_select(1, p1).contiguous())              # (m,n), (n,n)
    p2 = torch.argsort((R1 * R1).sum(1), descending=True)
    Q2, R2 = qr(R1.index_select(0, p2).t().contiguous())         # (n,n), (n,n)
    Ux, S, Vtx = jacobi(R2.t().contiguous(), k)                  # (n,k), (k), (k,n)
    Ux_un = torch.empty_like(Ux)
    Ux_un[p2] = Ux                                               # P2^T Ux
    U = mm_nn(Q1.contiguous(), Ux_un)                            # (m,k)
    Vt_p = mm_nt(Vtx, Q2.contiguous())                           # (k,n) = Vx^T Q2^T
    Vt = torch.empty_like(Vt_p)
    Vt[:, p1] = Vt_p                                             # undo the column sort
    return U, S, Vt


class CudaBackend:
    name = 'cuda'

    def __init__(self, device=None):
        self.lib = L.load()
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        sm, major, minor = C.c_int(), C.c_int(), C.c_int()
        with torch.cuda.device(self.device):
            L.check(self.lib.tn_device_info(C.byref(sm), C.byref(major), C.byref(minor)))
        self.sm_count = sm.value
        self._ws = {}
        self._scalars = torch.zeros(4096, dtype=torch.float64, device=self.device)
        self._scalar_pos = 0

    # ---- memory helpers ----
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def workspace(self, key, nbytes):
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    def from_numpy(self, x):
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=self.device)

    def to_numpy(self, t):
        return t.detach().cpu().numpy()

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def launch_count(self):
        return int(self.lib.tn_launch_count())

    # ---- a5: batched environment update ----
    def env_update(self, direction, T, outputs):
        """outputs: list over outgoing operators of lists of links (E or None, op (d,d) array or None).
        direction 0: left to right, 1: right to left.  Returns the list of new environment matrices."""
        a, d, b = T.shape
        T = T.contiguous()
        n_out = len(outputs)
        e_dim = b if direction == 0 else a
        outs = [self.empty(e_dim, e_dim) for _ in range(n_out)]
        begins, link_E, ops, has_op, keep = [0], [], [], [], []
        for links in outputs:
            for E, op in links:
                if E is not None:
                    E = E.contiguous()
                    keep.append(E)
                link_E.append(E.data_ptr() if E is not None else 0)
                has_op.append(0 if op is None else 1)
                ops.append(np.eye(d) if op is None else op)
            begins.append(len(link_E))
        n_links = len(link_E)
        nbytes = self.lib.tn_env_update_workspace_bytes(a, d, b, n_out, n_links)
        ws = self.workspace('env', nbytes)
        L.check(self.lib.tn_env_update(direction, _ptr(T), a, d, b, n_out,
                                       (C.c_void_p * n_out)(*[o.data_ptr() for o in outs]),
                                       (C.c_int * (n_out + 1))(*begins), (C.c_void_p * n_links)(*link_E),
                                       _op_array(ops, d), (C.c_int * n_links)(*has_op), _ptr(ws), ws.numel(),
                                       self.stream()))
        return outs

    def lincomb(self, xs, coeffs):
        out = torch.empty_like(xs[0])
        xs = [x.contiguous() for x in xs]
        n = len(xs)
        L.check(self.lib.tn_lincomb(_ptr(out), out.numel(), n, (C.c_void_p * n)(*[x.data_ptr() for x in xs]),
                                    (C.c_double * n)(*[float(c) for c in coeffs]), self.stream()))
        return out

    def site_op(self, T, op):
        a, d, b = T.shape
        T = T.contiguous()
        out = torch.empty_like(T)
        L.check(self.lib.tn_apply_site_op(_ptr(out), _ptr(T), a, d, b, _op_array([op], d), self.stream()))
        return out

    # ---- reductions to device scalars (read back in one copy by `scalars_to_host`) ----
    def _slot(self):
        if self._scalar_pos >= self._scalars.numel():
            raise L.TnError('scalar slots exhausted; call scalars_to_host()')
        s = self._scalars[self._scalar_pos:self._scalar_pos + 1]
        self._scalar_pos += 1
        return s

    def dot(self, x, y):
        """returns a device scalar slot holding sum(x*y)"""
        x, y = x.contiguous(), y.contiguous()
        slot = self._slot()
        nbytes = self.lib.tn_dot_workspace_bytes(x.numel())
        ws = self.workspace('dot', nbytes)
        L.check(self.lib.tn_dot(_ptr(x), _ptr(y), x.numel(), _ptr(slot), _ptr(ws), ws.numel(), self.stream()))
        return slot

    def trace(self, E):
        E = E.contiguous()
        slot = self._slot()
        L.check(self.lib.tn_trace(_ptr(E), E.shape[0], _ptr(slot), self.stream()))
        return slot

    def scalars_to_host(self, slots):
        vals = torch.cat(slots).cpu().numpy() if slots else np.zeros(0)
        self._scalar_pos = 0
        return vals

    # ---- a6: mode products through the chain GEMM ----
    def _gemm(self, mode, M, N, K, A, B, lda, ldb, Cout, deterministic=1):
        prob = (L.TnProblem * 1)()
        prob[0].C, prob[0].alpha, prob[0].link_begin, prob[0].link_count, prob[0].accumulate = Cout.data_ptr(), 1.0, 0, 1, 0
        link = (L.TnLink * 1)()
        link[0].A, link[0].B, link[0].has_op = A.data_ptr(), B.data_ptr(), 0
        nbytes = self.lib.tn_chain_gemm_workspace_bytes(1, 1)
        ws = self.workspace('gemm', nbytes)
        L.check(self.lib.tn_chain_gemm(mode, M, N, K, 1, lda, ldb, N, prob, 1, link, 1, deterministic, _ptr(ws), ws.numel(),
                                       self.stream()))
        return Cout

    def mode_product(self, T, mat, bond):
        """out[.., j, ..] = sum_i T[.., i, ..] mat[i, j] (absorb_matrix2tensor, TensorBasicModule.py:387-424)."""
        a, d, b = T.shape
        T, mat = T.contiguous(), mat.contiguous()
        if bond == 0:
            k = mat.shape[1]
            out = self.empty(k, d, b)
            return self._gemm(2, k, d * b, a, mat, T, k, d * b, out)        # mat^T . T   (TN)
        if bond == 2:
            k = mat.shape[1]
            out = self.empty(a, d, k)
            return self._gemm(0, a * d, k, b, T, mat, b, k, out)            # T . mat     (NN)
        return self.site_op(T, mat.T.cpu().numpy())

    # ---- a1/a2: effective Hamiltonian ----
    def effh_plan(self, shape, HL=None, HR=None, M=None, LS=(), ls_ops=(), RS=(), rs_ops=(), XL=(), XR=(), x_coeff=(),
                  rank=0, world=1):
        a, d, b = shape
        n_ls, n_rs, n_x = len(LS), len(RS), len(XL)
        keep = [t.contiguous() if t is not None else None for t in [HL, HR, *LS, *RS, *XL, *XR]]
        HLc, HRc = keep[0], keep[1]
        LSc, RSc = keep[2:2 + n_ls], keep[2 + n_ls:2 + n_ls + n_rs]
        XLc, XRc = keep[2 + n_ls + n_rs:2 + n_ls + n_rs + n_x], keep[2 + n_ls + n_rs + n_x:]
        nbytes = self.lib.tn_effh_plan_workspace_bytes(a, d, b, n_ls, n_rs, n_x)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        handle = C.c_void_p()

        def parr(ts):
            return (C.c_void_p * max(len(ts), 1))(*[t.data_ptr() for t in ts]) if ts else None

        Mp = _op_array([M], d) if M is not None else None
        L.check(self.lib.tn_effh_plan_create(
            C.byref(handle), a, d, b, _ptr(HLc), _ptr(HRc), Mp, n_ls, parr(LSc), _op_array(ls_ops, d) if n_ls else None,
            n_rs, parr(RSc), _op_array(rs_ops, d) if n_rs else None, n_x, parr(XLc), parr(XRc),
            (C.c_double * max(n_x, 1))(*[float(c) for c in x_coeff]) if n_x else None, rank, world, _ptr(ws), ws.numel(),
            self.stream()))
        return EffHPlan(self, handle, ws, (a, d, b), keep)

    # ---- a8: eigensolver ----
    def lanczos(self, plan, tau, v0, tol, ncv=20, max_restarts=1000, allreduce=None):
        """dominant eigenpair of (1 - tau*H_eff): (lambda, vector, n_matvec, residual, converged)."""
        v0 = v0.contiguous().reshape(-1)
        n = v0.numel()
        nbytes = self.lib.tn_lanczos_workspace_bytes(n, ncv)
        ws = self.workspace('lanczos', nbytes)
        out = torch.empty_like(v0)
        lam, resid, nmv = C.c_double(), C.c_double(), C.c_int()
        cb = L.ALLREDUCE_FN(allreduce) if allreduce is not None else C.cast(None, L.ALLREDUCE_FN)
        st = self.lib.tn_lanczos_lm1(plan._handle, float(tau), _ptr(v0), float(tol), int(ncv), int(max_restarts),
                                     C.byref(lam), _ptr(out), C.byref(nmv), C.byref(resid), cb, None, _ptr(ws),
                                     ws.numel(), self.stream())
        if st not in (0, -4):
            L.check(st)
        return lam.value, out, nmv.value, resid.value, st == 0

    # ---- a7/a11/a12: factorizations ----
    def svd(self, A, k_keep=None, precondition=True):
        """A (m,n) -> U (m,k), S (k), Vt (k,n) with the k largest singular triplets.
        One-sided Jacobi (tn_svd_jacobi) after two QR steps with norm-sorted columns (Drmac-Veselic preconditioning, see
        preconditioned_svd below): a raw graded matrix needs 20-40 sweeps and loses relative accuracy on the small values, one
        QR step gives ~11 sweeps, two give ~8; sorting the columns by norm before each QR (a cheap stand-in for column
        pivoting) brings real two-site wavefunctions from 11-13 sweeps to 6-8.  U and Vt are one chain-GEMM call each."""
        A = A.contiguous()
        m, n = A.shape
        k = min(m, n) if k_keep is None else min(k_keep, m, n)
        if m < n:  # work on the transpose: A^T = U' S V'^T  =>  A = V' S U'^T
            U2, S, Vt2 = self.svd(A.t().contiguous(), k_keep=k, precondition=precondition)
            return Vt2.t().contiguous(), S, U2.t().contiguous()
        if precondition and n > 1:
            if n >= 64:
                def mm_nn(X, Y):   # X (p,q) . Y (q,r)
                    out = self.empty(X.shape[0], Y.shape[1])
                    return self._gemm(0, X.shape[0], Y.shape[1], X.shape[1], X.contiguous(), Y.contiguous(), X.shape[1], Y.shape[1], out)

                def mm_nt(X, Y):   # X (p,q) . Y (r,q)^T
                    out = self.empty(X.shape[0], Y.shape[0])
                    return self._gemm(1, X.shape[0], Y.shape[0], X.shape[1], X.contiguous(), Y.contiguous(), X.shape[1], Y.shape[1], out)
                return preconditioned_svd(A, k, self.qr, self._jacobi, mm_nn, mm_nt)
            Q1, R1 = self.qr(A)                                  # (m,n), (n,n)
            Ux, S, Vtx = self._jacobi(R1.t().contiguous())       # R1^T = Ux S Vtx  =>  A = (Q1 Vtx^T) S Ux^T
            U = self.empty(m, n)
            self._gemm(1, m, n, n, Q1.contiguous(), Vtx, n, n, U)   # U[i,j] = sum_l Q1[i,l] Vtx[j,l]   (NT)
            Vt = Ux.t().contiguous()
            if k < n:
                U, S, Vt = U[:, :k].contiguous(), S[:k].contiguous(), Vt[:k].contiguous()
            return U, S, Vt
        return self._jacobi(A, k)

    def _jacobi(self, A, k_keep=None):
        A = A.contiguous()
        m, n = A.shape
        k = min(m, n) if k_keep is None else min(k_keep, m, n)
        U, S, Vt = self.empty(m, k), self.empty(k), self.empty(k, n)
        nbytes = self.lib.tn_svd_workspace_bytes(m, n)
        ws = self.workspace('svd', nbytes)
        sweeps = C.c_int()
        L.check(self.lib.tn_svd_jacobi(_ptr(A), m, n, k, _ptr(U), _ptr(S), _ptr(Vt), C.byref(sweeps), _ptr(ws), ws.numel(),
                                       self.stream()))
        self.last_svd_sweeps = sweeps.value
        return U, S, Vt

    def qr(self, A):
        """thin QR of gauge moves (np.linalg.qr at TensorBasicModule.py:342-345).  Round 1: cuSOLVER via torch."""
        return torch.linalg.qr(A, mode='reduced')

    def scale_diag_rows(self, S, Vt):
        return S[:, None] * Vt

    def norm(self, x):
        return float(torch.linalg.vector_norm(x))
