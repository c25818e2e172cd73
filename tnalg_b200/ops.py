"""Host-side operator layer: torch CUDA tensors in, C-ABI calls out.

`backend()` returns the process-wide CudaBackend.  Every method maps 1:1 onto an entry point of
include/tnalg_b200.h; torch is used for device memory, streams and the bootstrap of the multi-GPU communicator
(torch.distributed carries the 128-byte NCCL id once; every collective on the data path is issued by the library).
There is no library math on the path: QR, SVD, eigh, Lanczos and all contractions run in libtnalg_b200.so.
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
        rb, rc = C.c_int(), C.c_int()
        self.rows = be.lib.tn_effh_plan_rows(handle, C.byref(rb), C.byref(rc)) == 1
        self.row_begin, self.row_count = rb.value, rc.value

    def matvec(self, psi, c_id=0.0, c_h=1.0, out=None):
        """out = c_id*psi + c_h*H_eff psi; (c_id, c_h) = (1, -tau) is the reference handle (MPSClass.py:755-776)."""
        be = self._be
        psi = psi.contiguous()
        if out is None:   # a row-sliced plan reads the full psi and returns its (row_count, d, b) slice
            out = be.empty(self.row_count, *self.shape[1:]) if self.rows else torch.empty_like(psi)
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


class Comm:
    """Handle on the in-library communicator (tn_comm): rank, world and the stream-ordered collectives the hot path needs.
    Created once per process by CudaBackend.comm(); torch.distributed only carries the 128-byte NCCL id (bootstrap)."""

    def __init__(self, be, handle, rank, world):
        self._be, self._handle, self.rank, self.world = be, handle, rank, world

    def allreduce(self, t):
        L.check(self._be.lib.tn_comm_allreduce_sum(self._handle, _ptr(t), t.numel(), self._be.stream()))
        return t

    def broadcast(self, t, src=0):
        self.broadcast_many([t], [src])
        return t

    def broadcast_many(self, tensors, roots):
        """one grouped NCCL launch for all the outgoing operators of a sharded environment update"""
        n = len(tensors)
        if n == 0:
            return
        L.check(self._be.lib.tn_comm_broadcast_many(self._handle, (C.c_void_p * n)(*[t.data_ptr() for t in tensors]),
                                                    (C.c_longlong * n)(*[t.numel() for t in tensors]),
                                                    (C.c_int * n)(*[int(r) for r in roots]), n, self._be.stream()))

    def allgather_inplace(self, buf):
        """buf: (world * k, ...) contiguous; rank r has written rows [r*k, (r+1)*k); ONE all-gather fills in the others"""
        per = buf.numel() // self.world
        assert buf.is_contiguous() and per * self.world == buf.numel()
        send = buf.data_ptr() + 8 * per * self.rank
        L.check(self._be.lib.tn_comm_allgather(self._handle, C.c_void_p(send), _ptr(buf), per, self._be.stream()))
        return buf

    def collectives(self):
        return int(self._be.lib.tn_comm_collectives(self._handle))

    def peer_collectives(self):
        """collectives that ran as stores into the NVLink peer window instead of NCCL launches"""
        return int(self._be.lib.tn_comm_peer_collectives(self._handle))

    def destroy(self):
        if self._handle is not None:
            self._be.lib.tn_comm_destroy(self._handle)
            self._handle = None


class CudaBackend:
    name = 'cuda'
    # multi-GPU decomposition of a local eigenproblem: 'rows' = every rank computes a row slice of H|psi> for all terms and
    # keeps a slice of the Krylov basis (all-gather per step); 'terms' = the coupling-term links are dealt round-robin and the
    # partial H|psi> is all-reduced (the decomposition SURVEY.md 8e names)
    shard_mode = 'rows'
    shard_min_rows = 16      # a rank needs at least this many rows of the (a, d*b) output, else the site is solved replicated
    shard_min_n = 1 << 14    # ... and the vector at least this many elements

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
        self._comm = None

    # ---- multi-GPU communicator ----
    def comm(self):
        """the library communicator over the ranks of torch.distributed (None for a single process).  The NCCL unique id is
        created by rank 0 (tn_comm_unique_id) and broadcast once through torch.distributed; the communicator itself and every
        collective issued through it live in libtnalg_b200.so."""
        if self._comm is not None:
            return self._comm
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return None
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = C.create_string_buffer(128)
        if rank == 0:
            L.check(self.lib.tn_comm_unique_id(ident))
        box = [bytes(ident.raw)]
        dist.broadcast_object_list(box, src=0)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.tn_comm_init_rank(C.byref(handle), box[0], rank, world))
        self._comm = Comm(self, handle, rank, world)
        return self._comm

    def dmma_peak_tflops(self):
        """issue-rate ceiling of the FP64 tensor pipe measured on this device (tn_measure_dmma_peak)"""
        ws = self.workspace('peak', 8 * 512 * self.sm_count)
        out = C.c_double()
        L.check(self.lib.tn_measure_dmma_peak(C.byref(out), _ptr(ws), ws.numel(), self.stream()))
        return out.value

    def set_deterministic(self, on=True):
        """bit-reproducible kernels (no stream-K split, no FP64-atomic tile combination, no concurrent matvec stages)"""
        return bool(self.lib.tn_set_deterministic(1 if on else 0))

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

    def to_numpy_many(self, tensors):
        """device tensors -> numpy arrays through page-locked staging buffers: all copies are queued on the stream and waited for
        once.  The arrays are views of pinned host memory (torch's caching host allocator recycles it when they are dropped)."""
        host = []
        for t in tensors:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t.detach(), non_blocking=True)
            host.append(h)
        torch.cuda.current_stream().synchronize()
        return [h.numpy() for h in host]

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def launch_count(self):
        return int(self.lib.tn_launch_count())

    # ---- a5: batched environment update ----
    def env_update(self, direction, T, outputs, outs=None):
        """outputs: list over outgoing operators of lists of links (E or None, op (d,d) array or None).
        direction 0: left to right, 1: right to left.  Returns the list of new environment matrices (written into `outs`
        -- contiguous (e, e) tensors -- when given)."""
        a, d, b = T.shape
        T = T.contiguous()
        n_out = len(outputs)
        e_dim = b if direction == 0 else a
        if outs is None:
            outs = [self.empty(e_dim, e_dim) for _ in range(n_out)]
        assert len(outs) == n_out and all(o.is_contiguous() and o.shape == (e_dim, e_dim) for o in outs)
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
    def shard_rows(self, a, d, b, rank, world):
        """(row_begin, row_count) of this rank's slice of the (a, d*b) output, or None when the site is too small to slice
        (every rank then solves the whole problem; rows_per = ceil(a / world) is what tn_lanczos_lm1 expects)"""
        if world <= 1 or self.shard_mode != 'rows':
            return None
        rows_per = -(-a // world)
        last = a - rows_per * (world - 1)
        if last < min(self.shard_min_rows, rows_per) or rows_per < self.shard_min_rows or a * d * b < self.shard_min_n:
            return None
        return rank * rows_per, min(rows_per, a - rank * rows_per)

    def effh_plan(self, shape, HL=None, HR=None, M=None, LS=(), ls_ops=(), RS=(), rs_ops=(), XL=(), XR=(), x_coeff=(),
                  rank=0, world=1, rows=None):
        """rows = (row_begin, row_count): a row-sliced plan (tn_effh_plan_create_rows); otherwise rank/world deal the links
        round-robin (term sharding) and (0, 1) is the full operator"""
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
        create, tail = (self.lib.tn_effh_plan_create, (rank, world)) if rows is None else (self.lib.tn_effh_plan_create_rows, tuple(rows))
        L.check(create(
            C.byref(handle), a, d, b, _ptr(HLc), _ptr(HRc), Mp, n_ls, parr(LSc), _op_array(ls_ops, d) if n_ls else None,
            n_rs, parr(RSc), _op_array(rs_ops, d) if n_rs else None, n_x, parr(XLc), parr(XRc),
            (C.c_double * max(n_x, 1))(*[float(c) for c in x_coeff]) if n_x else None, int(tail[0]), int(tail[1]), _ptr(ws), ws.numel(),
            self.stream()))
        return EffHPlan(self, handle, ws, (a, d, b), keep)

    # ---- a8: eigensolver ----
    def lanczos(self, plan, tau, v0, tol, ncv=20, max_restarts=1000, comm=None, allreduce=None):
        """dominant eigenpair of (1 - tau*H_eff): (lambda, vector, n_matvec, residual, converged).
        comm: the library communicator when the plan is row-sliced or term-sharded (collectives issued by the library);
        allreduce: a host callback for term-sharded plans instead (kept for callers that bring their own collective)."""
        v0 = v0.contiguous().reshape(-1)
        n = v0.numel()
        ch = comm._handle if comm is not None else None
        nbytes = self.lib.tn_lanczos_workspace_bytes_sharded(plan._handle, ch, int(ncv)) if ch is not None else \
            self.lib.tn_lanczos_workspace_bytes(n, ncv)
        ws = self.workspace('lanczos', nbytes)
        out = torch.empty_like(v0)
        lam, resid, nmv = C.c_double(), C.c_double(), C.c_int()
        cb = L.ALLREDUCE_FN(allreduce) if allreduce is not None else C.cast(None, L.ALLREDUCE_FN)
        st = self.lib.tn_lanczos_lm1(plan._handle, float(tau), _ptr(v0), float(tol), int(ncv), int(max_restarts),
                                     C.byref(lam), _ptr(out), C.byref(nmv), C.byref(resid), cb, None, ch, _ptr(ws),
                                     ws.numel(), self.stream())
        if st not in (0, -4):   # -4 (iteration limit) still writes the best Ritz pair; everything else is an error
            L.check(st)
        return lam.value, out, nmv.value, resid.value, st == 0

    def lanczos_generic(self, matvec, n, tau, v0, tol, ncv=20, max_restarts=1000, locked=None):
        """the same device-resident solver for a caller-supplied operator: matvec(x, y) receives two float64 device tensors of
        length n (aliases of the Krylov workspace) and must write y = H x.  locked: (n_locked, n) tensor of orthonormal vectors
        the search stays orthogonal to (deflation)."""
        v0 = v0.contiguous().reshape(-1)
        nbytes = self.lib.tn_lanczos_workspace_bytes(n, ncv)
        ws = self.workspace('lanczos', nbytes)
        out = torch.empty_like(v0)
        lam, resid, nmv = C.c_double(), C.c_double(), C.c_int()
        err = []

        def cb(xp, yp, user, stream):
            try:
                matvec(_alias(xp, n, self.device), _alias(yp, n, self.device))
                return 0
            except Exception as e:  # the C side turns the status into TN_ERR_INVALID; keep the Python error for the caller
                err.append(e)
                return 1
        n_locked = 0 if locked is None else int(locked.shape[0])
        if n_locked:
            locked = locked.contiguous()
        st = self.lib.tn_lanczos_generic(L.MATVEC_FN(cb), None, n, float(tau), _ptr(v0), float(tol), int(ncv), int(max_restarts),
                                         _ptr(locked) if n_locked else None, n_locked, n if n_locked else 0, C.byref(lam), _ptr(out),
                                         C.byref(nmv), C.byref(resid), _ptr(ws), ws.numel(), self.stream())
        if err:
            raise err[0]
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
        """thin Householder QR (np.linalg.qr at TensorBasicModule.py:342-345): A (m,n) -> Q (m,k), R (k,n) -- tn_qr_householder"""
        A = A.contiguous()
        m, n = A.shape
        k = min(m, n)
        Q, R = self.empty(m, k), self.empty(k, n)
        ws = self.workspace('qr', self.lib.tn_qr_workspace_bytes(m, n))
        L.check(self.lib.tn_qr_householder(_ptr(A), m, n, 0, _ptr(Q), 0, _ptr(R), _ptr(ws), ws.numel(), self.stream()))
        return Q, R

    def qr_tensor(self, T, left2right):
        """the QR gauge move on a site tensor (a,d,b) without materialising a transposed matricisation:
        left2right: T.reshape(a*d, b) = Q R -> (Q (a,d,k), R (k,b));  else T.reshape(a, d*b)^T = Q R -> (Q^T (k,d,b), R (k,a))"""
        T = T.contiguous()
        a, d, b = T.shape
        if left2right:
            k = min(a * d, b)
            Q, R = self.empty(a, d, k), self.empty(k, b)
            ws = self.workspace('qr', self.lib.tn_qr_workspace_bytes(a * d, b))
            L.check(self.lib.tn_qr_l2r(_ptr(T), a, d, b, _ptr(Q), _ptr(R), _ptr(ws), ws.numel(), self.stream()))
        else:
            k = min(a, d * b)
            Q, R = self.empty(k, d, b), self.empty(k, a)
            ws = self.workspace('qr', self.lib.tn_qr_workspace_bytes(d * b, a))
            L.check(self.lib.tn_qr_r2l(_ptr(T), a, d, b, _ptr(Q), _ptr(R), _ptr(ws), ws.numel(), self.stream()))
        return Q, R

    def eigh(self, A):
        """symmetric eigenproblem on the Jacobi kernels (tn_eigh_jacobi): ascending eigenvalues, eigenvectors in columns"""
        A = A.contiguous()
        n = A.shape[0]
        w, V = self.empty(n), self.empty(n, n)
        ws = self.workspace('eigh', self.lib.tn_eigh_workspace_bytes(n))
        sweeps = C.c_int()
        L.check(self.lib.tn_eigh_jacobi(_ptr(A), n, _ptr(w), _ptr(V), C.byref(sweeps), _ptr(ws), ws.numel(), self.stream()))
        return w, V

    # ---- a10 through the C ABI: per-term chains (the batched, prefix-sharing form is envs.expect_products) ----
    def expect_terms(self, mps, center, terms):
        """terms: list of ((site, op (d,d) array),) or ((site1, op1), (site2, op2)) -> numpy array of expectation values
        (tn_expect_1body / tn_expect_2body)"""
        Ls, d = len(mps), mps[0].shape[1]
        ts = [t.contiguous() for t in mps]
        dims = (C.c_int * (Ls + 1))(*([ts[0].shape[0]] + [t.shape[2] for t in ts]))
        ptrs = (C.c_void_p * Ls)(*[t.data_ptr() for t in ts])
        out = np.zeros(len(terms))
        for nbody in (1, 2):
            idx = [i for i, tm in enumerate(terms) if len(tm) == nbody]
            if not idx:
                continue
            n = len(idx)
            ws = self.workspace('expect', self.lib.tn_expect_workspace_bytes(dims, Ls, d, n))
            res = (C.c_double * n)()
            if nbody == 1:
                st = self.lib.tn_expect_1body(ptrs, dims, Ls, d, int(center), n, (C.c_int * n)(*[int(terms[i][0][0]) for i in idx]),
                                              _op_array([terms[i][0][1] for i in idx], d), res, _ptr(ws), ws.numel(), self.stream())
            else:
                st = self.lib.tn_expect_2body(ptrs, dims, Ls, d, int(center), n, (C.c_int * n)(*[int(terms[i][0][0]) for i in idx]),
                                              (C.c_int * n)(*[int(terms[i][1][0]) for i in idx]),
                                              _op_array([terms[i][0][1] for i in idx], d), _op_array([terms[i][1][1] for i in idx], d),
                                              res, _ptr(ws), ws.numel(), self.stream())
            L.check(st)
            out[idx] = np.array(res[:])
        return out

    # ---- (f)4: exact diagonalisation on the full d^L space ----
    def _ed_args(self, couplings, hamilts, d):
        couplings = np.asarray(couplings, dtype=int).reshape(-1, 3)
        nt, nh = couplings.shape[0], len(hamilts)
        hs = np.concatenate([np.asarray(np.real(h), dtype=float).reshape(-1) for h in hamilts])
        if hs.size != nh * d ** 4:
            raise ValueError('two-site Hamiltonians must be (d^2, d^2)')
        arr = lambda col: (C.c_int * nt)(*[int(x) for x in couplings[:, col]])  # noqa: E731
        return nt, nh, arr(0), arr(1), arr(2), (C.c_double * hs.size)(*hs)

    def ed_apply(self, v, L_sites, d, couplings, hamilts, c_id=1.0, c_h=-1e-4):
        """EDbasic.project_all_hamilt (library/EDspinClass.py:69-77): c_id*v + c_h*sum_n h[c_n](p1_n, p2_n) v"""
        v = v.contiguous().reshape(-1)
        nt, nh, p1, p2, hi, hs = self._ed_args(couplings, hamilts, d)
        out = torch.empty_like(v)
        ws = self.workspace('ed', self.lib.tn_ed_workspace_bytes(L_sites, d, nt, nh, 2))
        L.check(self.lib.tn_ed_apply(_ptr(out), _ptr(v), L_sites, d, nt, p1, p2, hi, hs, nh, float(c_id), float(c_h), _ptr(ws),
                                     ws.numel(), self.stream()))
        return out

    def ed_ground_state(self, v0, L_sites, d, couplings, hamilts, tau=1e-4, tol=1e-12, ncv=20, max_restarts=2000):
        """exact_ground_state (algorithms/ExactDiagonalizationAlgo.py:12-24): dominant eigenpair of 1 - tau*H on d^L"""
        v0 = v0.contiguous().reshape(-1)
        nt, nh, p1, p2, hi, hs = self._ed_args(couplings, hamilts, d)
        out = torch.empty_like(v0)
        ws = self.workspace('ed', self.lib.tn_ed_workspace_bytes(L_sites, d, nt, nh, ncv))
        lam, resid, nmv = C.c_double(), C.c_double(), C.c_int()
        st = self.lib.tn_ed_ground_state(L_sites, d, nt, p1, p2, hi, hs, nh, float(tau), _ptr(v0), float(tol), int(ncv),
                                         int(max_restarts), C.byref(lam), _ptr(out), C.byref(nmv), C.byref(resid), _ptr(ws),
                                         ws.numel(), self.stream())
        if st not in (0, -4):
            L.check(st)
        return lam.value, out, nmv.value, resid.value, st == 0

    def scale_diag_rows(self, S, Vt):
        return S[:, None] * Vt

    def norm(self, x):
        """2-norm through the deterministic dot kernel (tn_dot); blocking"""
        x = x.contiguous().reshape(-1)
        if not hasattr(self, '_norm_slot'):
            self._norm_slot = torch.zeros(1, dtype=torch.float64, device=self.device)
        ws = self.workspace('dot', self.lib.tn_dot_workspace_bytes(x.numel()))
        L.check(self.lib.tn_dot(_ptr(x), _ptr(x), x.numel(), _ptr(self._norm_slot), _ptr(ws), ws.numel(), self.stream()))
        return float(self._norm_slot.item()) ** 0.5


def _alias(ptr, count, device):
    """torch view of `count` float64 values at device address `ptr` (operands handed to a matvec callback)"""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {'shape': (int(count),), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}
    return torch.as_tensor(h, device=device)
