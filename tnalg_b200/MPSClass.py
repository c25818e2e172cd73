"""Drop-in replacement of the reference's MPSClass.MpsOpenBoundaryClass for the finite-size DMRG path.

Same constructor signature, attribute set, method names and pickled layout as MPSClass.py:26-1098 of the reference
(root-level generation); the numpy/scipy arithmetic underneath is replaced by the CUDA library through
tnalg_b200.ops (C ABI in include/tnalg_b200.h).  The MPS tensors live on the GPU for the whole run and are turned
into numpy arrays by clean_to_save() / pickling, so `.pr` files written through BasicFunctionsSJR.save_pr keep the
reference layout.

Deliberate differences:
  * effective operators are kept as summed complementary blocks per bond (tnalg_b200.envs) instead of one matrix
    per term per bond; `effect_s`, `effect_ss`, `pos_effect_*`, `effective_id`, `opt_env` exist for layout
    compatibility and stay empty;
  * observables are always computed from the actual tensors, so corr_* after calculate_entanglement_spectrum() are
    the correct values, not the reference's stale-cache values;
  * the eigensolver is a device-resident thick-restart Lanczos with ARPACK's convergence criterion;
  * is_parallel / is_env_parallel_lmr / par_pool are accepted and ignored (the GPU batches terms itself).
"""
import numpy as np

from . import ops as _ops
from .envs import EnvCache, TermTable, expect_products
from .HamiltonianModule import spin_operators

VERSION = '2018-06-3'  # MPSClass.py:15, kept so that pickles look alike


class _TensorList(list):
    """list of site tensors that converts assigned numpy arrays to device tensors and tells the owner which site
    changed (so that cached environments are invalidated)."""

    def __init__(self, owner, items):
        super().__init__(items)
        self._owner = owner

    def __setitem__(self, n, value):
        owner = self._owner
        if not hasattr(value, 'data_ptr'):
            value = owner._be.from_numpy(np.real(value))
        super().__setitem__(n, value)
        owner._tensor_changed(n)

    def __reduce__(self):
        return (list, (list(self),))


class MpsBasic:
    def __init__(self):
        self.version = VERSION
        self.operators = list()

    def append_operators(self, op_new):
        if type(op_new) is np.ndarray:
            self.operators.append(op_new)
        else:
            for n in range(0, len(op_new)):
                self.operators.append(op_new[n])


_STATE_KEYS = ('version', 'operators', 'spin', 'phys_dim', 'decomp_way', 'length', 'orthogonality', 'center', 'lm',
               'ent', 'mps', 'virtual_dim', '_is_save_op', 'effect_s', 'pos_effect_s', 'effect_ss', 'pos_effect_ss',
               'effective_id', 'opt_env', '_is_parallel', 'pool', '_debug', 'eig_way', '_is_env_parallel_lmr')


class MpsOpenBoundaryClass(MpsBasic):
    """Open-boundary MPS with the reference's interface (MPSClass.py:53-55):
    MpsOpenBoundaryClass(length, d, chi, spin='half', way='qr', ini_way='r', operators=None, debug=False,
                         is_parallel=False, is_save_op=False, eig_way=0, par_pool=None, is_env_parallel_lmr=True)"""

    def __init__(self, length, d, chi, spin='half', way='qr', ini_way='r', operators=None, debug=False,
                 is_parallel=False, is_save_op=False, eig_way=0, par_pool=None, is_env_parallel_lmr=True):
        MpsBasic.__init__(self)
        self.spin = spin
        self.phys_dim = d
        self.decomp_way = way
        self.length = length
        self.orthogonality = np.zeros((length, 1))
        self.center = -1
        self.lm = [np.zeros(0) for _ in range(length - 1)]
        self.ent = np.zeros((self.length - 1, 1))
        self._be = _ops.backend()
        if ini_way == 'r':
            # same np.random.randn draw order as random_open_mps (TensorBasicModule.py:181-186)
            host = [None] * length
            host[0] = np.random.randn(1, d, chi)
            host[length - 1] = np.random.randn(chi, d, 1)
            for n in range(1, length - 1):
                host[n] = np.random.randn(chi, d, chi)
        elif ini_way == '1':
            host = [np.ones((1, d, chi))] + [np.ones((chi, d, chi)) for _ in range(length - 2)] + [np.ones((chi, d, 1))]
        else:
            raise ValueError("ini_way must be 'r' or '1'")
        self.mps = _TensorList(self, [self._be.from_numpy(t) for t in host])
        self.virtual_dim = np.ones((length + 1,)).astype(int) * chi
        self.virtual_dim[0] = 1
        self.virtual_dim[-1] = 1
        if operators is None:
            op_half = spin_operators(spin)
            self.operators = [op_half['id'], op_half['sx'], op_half['sy'], op_half['sz'], op_half['su'], op_half['sd']]
        else:
            self.operators = operators
        self._is_save_op = is_save_op
        self.effect_s = {'none': np.zeros(0)}
        self.pos_effect_s = np.zeros((0, 3)).astype(int)
        self.effect_ss = {'none': np.zeros(0)}
        self.pos_effect_ss = np.zeros((0, 5)).astype(int)
        self.effective_id = {'none': np.zeros(0)}
        self.opt_env = dict()
        self._is_parallel = is_parallel
        self.pool = None
        self._debug = debug
        self.eig_way = eig_way
        self._is_env_parallel_lmr = is_env_parallel_lmr
        self._init_runtime()

    # ---- runtime (non-pickled) state ----
    def _init_runtime(self):
        self._env = None
        self._env_key = None
        self.stats = {'n_solves': 0, 'n_matvec': 0, 'flops_algorithmic': 0.0, 'flops_executed': 0.0,
                      't_solve': 0.0, 'not_converged': 0}
        self.lanczos_ncv = 20           # ARPACK's default ncv for k=1 (scipy eigsh)
        self.lanczos_max_restarts = 2000
        self.timing = False             # when True, update_tensor_eigs records per-phase times with CUDA events
        self._events = []               # solver (Lanczos) intervals
        self._phase_events = {'gauge': [], 'env': []}

    def _ensure_device(self):
        if not hasattr(self, '_be') or self._be is None:
            self._be = _ops.backend()
        if not hasattr(self, '_env'):
            self._init_runtime()
        if not isinstance(self.mps, _TensorList):
            self.mps = _TensorList(self, [t.to(self._be.device) if hasattr(t, 'data_ptr') else self._be.from_numpy(np.real(t))
                                          for t in self.mps])
            self._env = None

    def load_tensors(self, tensors, center, virtual_dim=None):
        """replace the site tensors by `tensors` (numpy arrays or torch tensors, e.g. pinned host memory or the `mps`
        list of a revived `.pr` pickle), copy them to the device and declare `center` the orthogonality centre."""
        be = self._be = _ops.backend()
        if not hasattr(self, '_env'):
            self._init_runtime()
        dev = []
        for t in tensors:
            if hasattr(t, 'data_ptr'):
                dev.append(t.to(be.device, non_blocking=True))
            else:
                dev.append(be.from_numpy(np.real(t)))
        if len(dev) != self.length:
            raise ValueError('expected %d site tensors, got %d' % (self.length, len(dev)))
        self.mps = _TensorList(self, dev)
        self.virtual_dim = np.array([dev[0].shape[0]] + [t.shape[2] for t in dev]) if virtual_dim is None else virtual_dim
        self.center = center
        self.orthogonality = np.zeros((self.length, 1))
        self.orthogonality[:center] = -1
        self.orthogonality[center + 1:] = 1
        self._env = None
        self._env_key = None

    def _tensor_changed(self, n):
        env = getattr(self, '_env', None)
        if env is not None:
            env.invalidate_site(n)

    def __getstate__(self):
        state = {}
        for k in _STATE_KEYS:
            v = getattr(self, k)
            if k == 'mps':
                v = [self._be.to_numpy(t) if hasattr(t, 'data_ptr') else np.asarray(t) for t in v]
            state[k] = v
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

    # ---- gauge moves (a7: orthogonalize_mps / correct_orthogonal_center, MPSClass.py:143-198) ----
    def _decompose(self, n, left2right):
        """left2right/right2left_decompose_tensor (TensorBasicModule.py:314-384): returns (Q tensor, R, dim, lm) with
        T = Q R (left2right, R is (k,b)) or T = R^T-absorbed form (right2left)."""
        be = self._be
        T = self.mps[n]
        a, d, b = T.shape
        if not (self.decomp_way == 1 or self.decomp_way == 'svd'):
            Qt, R = be.qr_tensor(T, left2right)          # tn_qr_l2r / tn_qr_r2l (Householder, own kernels)
            return Qt, R, R.shape[0], np.zeros(0)
        mat = T.reshape(a * d, b) if left2right else T.reshape(a, d * b).t().contiguous()
        k = min(mat.shape)
        U, S, Vt = be.svd(mat)
        R = be.scale_diag_rows(S, Vt)
        lm = be.to_numpy(S)
        Q = U
        if lm.size and lm.min() <= 1e-14 * lm.max() * max(mat.shape):
            # numerically rank-deficient (e.g. ini_way='1'): the Jacobi kernel returns zero columns of U for zero singular
            # values, np.linalg.svd an orthonormal completion.  One Householder QR of U completes the basis, U = Q2 R2 with
            # R2 = diag(+-1) on the kept directions and 0 elsewhere, so T = Q2 (R2 S Vt) is unchanged and Q2 is an isometry.
            Q, R2 = be.qr(U.contiguous())
            R = be.mode_product(R.contiguous().reshape(k, 1, -1), R2.t().contiguous(), 0).reshape(k, -1)   # R2 . (S Vt)
        if left2right:
            Qt = Q.contiguous().reshape(a, d, k)
        else:
            Qt = Q.t().contiguous().reshape(k, d, b)
        return Qt, R, k, lm

    def orthogonalize_mps(self, l0, l1, normalize=False, is_trun=False, chi=-1):
        """gauge moves from l0 to l1 (MPSClass.py:143-167).  The extra arguments are those of the library generation
        (library/MPSClass.py:186-247): with is_trun the decomposition is an SVD and the new bond keeps the chi largest
        singular triplets (tn_svd_jacobi with k_keep); normalize rescales the tensor that absorbed the remainder."""
        self._ensure_device()
        be = self._be
        way0 = self.decomp_way
        if is_trun:
            self.decomp_way = 'svd'
        try:
            step = 1 if l0 < l1 else -1
            for n in range(l0, l1, step):
                Q, R, dim, lm = self._decompose(n, step == 1)
                if is_trun and dim > chi > 0:
                    # keep the leading chi triplets: Q columns / rows and the matching rows of R = diag(lm) Vt
                    Q = Q[:, :, :chi].contiguous() if step == 1 else Q[:chi].contiguous()
                    R = R[:chi].contiguous()
                    lm = lm[:chi]
                    dim = chi
                bond = n + 1 if step == 1 else n
                self.virtual_dim[bond] = dim
                if lm.size > 0 and self.center > -1:
                    self.lm[bond - 1] = lm.copy()
                self.mps[n] = Q
                # absorb_matrix2tensor(neighbour, R^T, 0 or 2)
                self.mps[n + step] = be.mode_product(self.mps[n + step], R.t().contiguous(), 0 if step == 1 else 2)
                if normalize:
                    self.mps[n + step] = self.mps[n + step] / be.norm(self.mps[n + step])
            if l0 < l1:
                self.orthogonality[l0:l1] = -1
                self.orthogonality[l1] = 0
            elif l0 > l1:
                self.orthogonality[l0:l1:-1] = 1
                self.orthogonality[l1] = 0
        finally:
            self.decomp_way = way0

    def truncate_virtual_bonds(self, chi1, center, way='full'):
        """reduce every bond to at most chi1 (library/MPSClass.py:909-923).  way='full': SVD-truncating sweep from site 0 to
        L-1 (optimal bond by bond); way='simple': cut the index ranges and re-orthogonalise."""
        self._ensure_device()
        if way == 'simple':
            for n in range(self.length):
                t = self.mps[n]
                self.mps[n] = t[:min(t.shape[0], chi1), :, :min(t.shape[2], chi1)].contiguous()
            self.virtual_dim = np.array([self.mps[0].shape[0]] + [t.shape[2] for t in self.mps])
            self.center = -1
            self.central_orthogonalization(center)
            self.mps[center] = self.mps[center] / self._be.norm(self.mps[center])
        else:
            self.correct_orthogonal_center(0)
            self.orthogonalize_mps(0, self.length - 1, normalize=True, is_trun=True, chi=chi1)
            self.center = self.length - 1
            if center != self.length - 1:
                self.correct_orthogonal_center(center)

    def central_orthogonalization(self, lc, l0=0, l1=-1):
        if l1 == -1:
            l1 = self.length - 1
        self.orthogonalize_mps(l0, lc)
        self.orthogonalize_mps(l1, lc)
        self.center = lc

    def correct_orthogonal_center(self, p=-1):
        if p < -0.5 and self.center < -0.5:
            p = self.check_orthogonal_center(if_print=False)
        elif p < -0.5:
            p = self.center
        if self.center < -0.5:
            self.central_orthogonalization(p)
        elif self.center != p:
            self.orthogonalize_mps(self.center, p)
        self.center = p

    def check_orthogonal_center(self, expected_center=-2, if_print=True):
        """recommend a centre from self.orthogonality (MPSClass.py:989-1024), no tensor is touched"""
        if self.center > -0.5:
            return self.center
        left = np.nonzero(self.orthogonality.reshape(-1) == -1)[0]
        return int(left[-1]) + 1 if left.size else 0

    # ---- a1/a2/a8: the local update (update_tensor_eigs, MPSClass.py:778-809) ----
    def _environments(self, index1, index2, coeff1, coeff2, tol):
        # keyed on CONTENT (the reference rebuilds its environments from coeff* on every call, MPSClass.py:633-682): an
        # in-place edit of the coupling arrays or of an operator (parameter ramps that reuse the dict) must not be missed
        import hashlib
        h = hashlib.blake2b(digest_size=16)
        for arr in (index1, index2, coeff1, coeff2):
            h.update(np.ascontiguousarray(np.asarray(arr, dtype=float)).tobytes())
            h.update(b'|')
        for o in self.operators:
            h.update(np.ascontiguousarray(np.asarray(o, dtype=complex)).tobytes())
        key = (h.digest(), float(tol))
        if self._env is None or self._env_key != key:
            terms = TermTable(index1, index2, coeff1, coeff2, self.operators, tol)
            if terms.length > self.length:
                raise ValueError('coupling terms reference site %d but the MPS has %d sites' % (terms.length - 1, self.length))
            self._env = EnvCache(self._be, terms, self.length)
            self._env.comm = self._comm()
            self._env_key = key
        return self._env

    def _comm(self):
        """the backend's communicator when the local problems of this MPS are sharded over the ranks (multi-GPU)"""
        if not getattr(self, 'shard_terms', True):
            return None
        return self._be.comm()

    def sync_replicas(self, src=0):
        """make every rank hold rank `src`'s tensors (one grouped broadcast).  The sharded path assumes bit-identical
        replicas; dmrg_finite_size calls this after drawing the random initial state, so ranks need not share a seed."""
        comm = self._comm()
        if comm is not None:
            self._ensure_device()
            comm.broadcast_many(list(self.mps), [src] * self.length)
            if self._env is not None:
                self._env.invalidate_all()

    def effective_hamiltonian_plan(self, p, index1, index2, coeff1, coeff2, tol=1e-12, rank=0, world=1, rows=None):
        """tn_effh_plan for site p (the opt_env groups of all_environments_optimized, MPSClass.py:633-682)."""
        self._ensure_device()
        env = self._environments(index1, index2, coeff1, coeff2, tol)
        return env.plan(p, self.mps, rank=rank, world=world, rows=rows)

    def _solve(self, make_plan, shape, v0, tau, tol):
        """dominant eigenvector of 1 - tau*H_eff for the local problem whose plan `make_plan(rank, world, rows)` builds.
        Multi-GPU: row-sliced plans + sliced Krylov basis ('rows'), or term-sharded plans + all-reduce ('terms'); sites too
        small to slice are solved by every rank and re-synchronised from rank 0.  All collectives are issued by the library."""
        import torch
        be = self._be
        comm = self._comm()
        a, d, b = shape
        rows, plan_comm, resync = None, None, False
        if comm is None:
            plan = make_plan(0, 1, None)
        else:
            mode = getattr(be, 'shard_mode', 'terms')
            rows = be.shard_rows(a, d, b, comm.rank, comm.world) if mode == 'rows' else None
            if rows is not None:
                plan, plan_comm = make_plan(0, 1, rows), comm
            elif mode == 'terms':
                plan, plan_comm = make_plan(comm.rank, comm.world, None), comm
            else:
                plan, resync = make_plan(0, 1, None), True
        n = a * d * b
        if self.timing:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.eig_way == 0 and comm is None:
            lam, vec, n_mv, resid, ok = self._dense_solve(plan, shape, v0, tau)
        else:
            lam, vec, n_mv, resid, ok = be.lanczos(plan, tau, v0.reshape(-1), tol, ncv=self.lanczos_ncv,
                                                   max_restarts=self.lanczos_max_restarts, comm=plan_comm)
        if self.timing:
            e1.record()
            self._events.append((e0, e1))
        if resync or (plan_comm is not None and rows is None):
            comm.broadcast(vec, src=0)   # replicated or all-reduced solves: keep the replicas bit-identical
        self.stats['n_solves'] += 1
        self.stats['n_matvec'] += n_mv
        self.stats['flops_algorithmic'] += n_mv * plan.flops_algorithmic   # whole job (all ranks together)
        self.stats['flops_executed'] += n_mv * plan.flops_executed         # this rank
        self.stats['not_converged'] += 0 if ok else 1
        self.last_eig = {'lambda': lam, 'residual': resid, 'n_matvec': n_mv, 'converged': ok, 'n': n,
                         'sharding': 'none' if comm is None else ('rows' if rows is not None else ('terms' if plan_comm else 'replicated'))}
        plan.destroy()
        return vec

    def _dense_solve(self, plan, shape, v0, tau):
        """eig_way = 0 (MPSClass.py:792-794): the explicit matrix 1 - tau*H_eff and its dominant eigenvector, here by the
        Jacobi eigensolver (tn_eigh_jacobi).  Small local problems only, as in the reference (O(n^2) memory)."""
        import torch
        be = self._be
        n = int(np.prod(shape))
        if n > 4096:
            raise ValueError('eig_way=0 builds the dense (n, n) effective Hamiltonian; n = %d is too large (limit 4096), use eig_way=1' % n)
        eye = torch.eye(n, dtype=torch.float64, device=be.device)
        h = be.empty(n, n)
        for j in range(n):
            plan.matvec(eye[j].reshape(shape), 1.0, -float(tau), out=h[j].reshape(shape))   # row j = column j (symmetric)
        h = 0.5 * (h + h.t())
        w, V = be.eigh(h.contiguous())
        w_host = be.to_numpy(w)
        best = int(np.argmax(np.abs(w_host)))
        vec = V[:, best].contiguous()
        if float((vec * v0.reshape(-1)).sum()) < 0:
            vec = -vec
        return float(w_host[best]), vec, n, 0.0, True

    def _phase(self, name):
        """context manager: CUDA-event interval of one phase of a local update (only while self.timing)"""
        import contextlib
        import torch

        @contextlib.contextmanager
        def cm():
            if not self.timing:
                yield
                return
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            try:
                yield
            finally:
                e1.record()
                self._phase_events[name].append((e0, e1))
        return cm()

    def update_tensor_eigs(self, p, index1, index2, coeff1, coeff2, tau, is_real, tol=1e-16):
        self._ensure_device()
        if self.center < -0.5:
            raise RuntimeError('CenterError: central-orthogonalize MPS before updating the tensor')
        with self._phase('gauge'):
            self.correct_orthogonal_center(p)
        env = self._environments(index1, index2, coeff1, coeff2, tol)
        shape = tuple(self.mps[p].shape)
        with self._phase('env'):
            env.ensure(p, self.mps)
        vec = self._solve(lambda rank, world, rows: env.plan(p, self.mps, rank=rank, world=world, rows=rows), shape,
                          self.mps[p], tau, tol)
        self.mps[p] = vec.reshape(shape)
        if self.eig_way == 1:
            self.opt_env = dict()

    # ---- two-site update with SVD truncation (north_star kernels 1 + 3; library/MPSClass.py:1676-1707 is the reference's
    # two-site machinery, used there by iDMRG; here it drives a finite-size sweep with bond-dimension growth) ----
    def update_two_sites_eigs(self, p, index1, index2, coeff1, coeff2, tau, is_real, tol=1e-16, chi=None, to_right=True):
        """optimise theta = mps[p] . mps[p+1] as the dominant eigenvector of 1 - tau*H_eff(two sites) and split it back with
        an SVD truncated to chi.  The centre ends on p+1 (to_right) or p."""
        self._ensure_device()
        be = self._be
        if self.center < -0.5:
            raise RuntimeError('CenterError: central-orthogonalize MPS before updating the tensor')
        if not 0 <= p < self.length - 1:
            raise ValueError('two-site update needs 0 <= p < length-1')
        from ._lib import MAX_PHYS_D
        if self.phys_dim ** 2 > MAX_PHYS_D:
            raise ValueError('two-site update: combined physical dimension d*d = %d exceeds TN_MAX_PHYS_DIM = %d'
                             % (self.phys_dim ** 2, MAX_PHYS_D))
        self.correct_orthogonal_center(p)
        env = self._environments(index1, index2, coeff1, coeff2, tol)
        a, d, k = self.mps[p].shape
        b = self.mps[p + 1].shape[2]
        chi = self.virtual_dim.max() if chi is None else chi
        # theta[(a s1), (s2 b)] = sum_k T_p[(a s1), k] T_{p+1}[k, (s2 b)]
        theta = be.mode_product(self.mps[p].reshape(a * d, 1, k), self.mps[p + 1].reshape(k, d * b), 2)
        vec = self._solve(lambda rank, world, rows: env.plan_two_site(p, self.mps, rank=rank, world=world, rows=rows),
                          (a, d * d, b), theta, tau, tol)
        kk = int(min(chi, a * d, d * b))
        U, S, Vt = be.svd(vec.reshape(a * d, d * b), k_keep=kk)          # Jacobi SVD, truncated to chi
        s_norm = be.norm(S)
        self.last_eig['truncation_error'] = max(0.0, 1.0 - s_norm ** 2)   # discarded weight (the eigenvector has norm 1)
        S = S / s_norm
        if to_right:
            self.mps[p] = U.contiguous().reshape(a, d, kk)
            self.mps[p + 1] = be.scale_diag_rows(S, Vt).contiguous().reshape(kk, d, b)
            self.center = p + 1
            self.orthogonality[p], self.orthogonality[p + 1] = -1, 0
        else:
            self.mps[p] = (U * S[None, :]).contiguous().reshape(a, d, kk)
            self.mps[p + 1] = Vt.contiguous().reshape(kk, d, b)
            self.center = p
            self.orthogonality[p], self.orthogonality[p + 1] = 0, 1
        self.virtual_dim[p + 1] = kk
        self.lm[p] = be.to_numpy(S)

    def solver_time_ms(self):
        """sum of the CUDA-event durations recorded by update_tensor_eigs while self.timing was True"""
        import torch
        torch.cuda.synchronize()
        t = sum(e0.elapsed_time(e1) for e0, e1 in self._events)
        self._events = []
        return t

    def phase_times_ms(self):
        """{'gauge': ms, 'env': ms} of the intervals recorded while self.timing was True (gauge = QR centre moves,
        env = environment updates of the bond moves)"""
        import torch
        torch.cuda.synchronize()
        out = {k: sum(e0.elapsed_time(e1) for e0, e1 in v) for k, v in self._phase_events.items()}
        self._phase_events = {k: [] for k in self._phase_events}
        return out

    # ---- a11: entanglement (MPSClass.py:812-839) ----
    def calculate_entanglement_spectrum(self, if_fast=True):
        self._ensure_device()
        _way, _center = self.decomp_way, self.center
        self.decomp_way = 'svd'
        if if_fast and _center > -0.5:
            p0, p1 = self.length - 1, 0
            for n in range(0, self.length - 1):
                if self.lm[n].size == 0:
                    p0, p1 = min(p0, n), max(p1, n)
            self.correct_orthogonal_center(p0)
            self.correct_orthogonal_center(p1 + 1)
            self.correct_orthogonal_center(_center)
        else:
            self.correct_orthogonal_center(0)
            self.correct_orthogonal_center(self.length - 1)
            if _center > 0:
                self.correct_orthogonal_center(_center)
        self.decomp_way = _way

    def calculate_entanglement_entropy(self):
        from .TensorBasicModule import entanglement_entropy
        for i in range(0, self.length - 1):
            self.ent[i] = -1 if self.lm[i].size == 0 else entanglement_entropy(self.lm[i])

    # ---- a10: observables (MPSClass.py:857-952) ----
    def _expect(self, terms):
        self._ensure_device()
        if self.center < -0.5:
            raise RuntimeError('observables need a centre-orthogonal MPS; call correct_orthogonal_center first')
        ops = [np.real(np.asarray(o)).astype(float) if np.abs(np.imag(np.asarray(o))).max() == 0 else None
               for o in self.operators]
        for term in terms:
            for _, sn in term:
                if ops[int(sn)] is None:
                    raise ValueError('operator %d is complex (e.g. sy); the real (is_real) path of this implementation cannot '
                                     'measure it -- see DESIGN.md, out of scope' % int(sn))
        out = []
        for i in range(0, len(terms), 1024):
            out.append(expect_products(self._be, self.mps, self.center, ops, terms[i:i + 1024], comm=self._comm()))
        return np.concatenate(out) if out else np.zeros(0)

    def observation_s1(self, inputs):
        sn, position = inputs
        return self._expect([((position, sn),)])[0]

    def observation_s1_s2(self, inputs):
        ssn, positions = inputs
        pair = sorted(zip([int(positions[0]), int(positions[1])], [int(ssn[0]), int(ssn[1])]))
        return self._expect([tuple(pair)])[0]

    def observe_magnetization(self, sn):
        return self._expect([((i, sn),) for i in range(self.length)]).reshape(-1, 1)

    def observe_bond_energy(self, index2, coeff2):
        index2 = np.asarray(index2, dtype=int)
        terms = [tuple(sorted(((int(r[0]), int(r[2])), (int(r[1]), int(r[3]))))) for r in index2]
        vals = self._expect(terms)
        return (np.asarray(coeff2, dtype=float).reshape(-1) * vals).reshape(-1, 1)

    def observe_bond_energy_and_magnetization(self, index2, coeff2, sns=(1, 3)):
        """observe_bond_energy + observe_magnetization(sn) for sn in sns in ONE pass over the chain (the density chain from the
        orthogonality centre and the operator-carrying environments are shared): what DMRG_anyH.py:57-63 evaluates after a sweep"""
        index2 = np.asarray(index2, dtype=int)
        terms = [tuple(sorted(((int(r[0]), int(r[2])), (int(r[1]), int(r[3]))))) for r in index2]
        nb = len(terms)
        for sn in sns:
            terms += [((i, int(sn)),) for i in range(self.length)]
        vals = self._expect(terms)
        eb = (np.asarray(coeff2, dtype=float).reshape(-1) * vals[:nb]).reshape(-1, 1)
        mags = [vals[nb + k * self.length: nb + (k + 1) * self.length].reshape(-1, 1) for k in range(len(sns))]
        return eb, mags

    def observe_correlators_from_middle(self, op1, op2, ob_len=None):
        if ob_len is None:
            ob_len = self.length
        pairs = []
        pos_mid = round(self.length / 2)
        pos1, pos2, n_control = pos_mid, pos_mid + 1, 0
        while pos1 > -0.1 and pos2 < self.length and ob_len > -0.1:
            pairs.append(((pos1, op1), (pos2, op2)))
            if n_control % 2 == 0:
                pos1 -= 1
            else:
                pos2 += 1
            n_control += 1
            ob_len -= 1
        return self._expect(pairs)

    def norm_mps(self):
        self._ensure_device()
        if self.center < -0.5:
            raise RuntimeError('norm_mps needs a centre-orthogonal MPS')
        return self._be.norm(self.mps[self.center])

    def reduced_density_matrix_two_body(self, p1, p2):
        """rho[(s1 s2), (s1' s2')] of sites p1 < p2, trace 1 (MPSClass.py:841-855).  Computed as the d^4 expectation values
        of matrix-unit products |s1><s1'| (x) |s2><s2'| in one batched pass of the observable machinery (any centre)."""
        p1, p2 = int(p1), int(p2)
        if not 0 <= p1 < p2 < self.length:
            raise ValueError('reduced_density_matrix_two_body needs 0 <= p1 < p2 < length')
        self._ensure_device()
        if self.center < -0.5:
            raise RuntimeError('observables need a centre-orthogonal MPS; call correct_orthogonal_center first')
        d = self.phys_dim
        units = []
        for a in range(d):
            for b in range(d):
                u = np.zeros((d, d))
                u[a, b] = 1.0
                units.append(u)
        terms = [((p1, a * d + ap), (p2, b * d + bp)) for a in range(d) for b in range(d) for ap in range(d) for bp in range(d)]
        vals = []
        for i in range(0, len(terms), 1024):
            vals.append(expect_products(self._be, self.mps, self.center, units, terms[i:i + 1024]))
        rho = np.concatenate(vals).reshape(d * d, d * d)      # [(a b), (a' b')] = <|a><a'| (x) |b><b'|>
        return rho / np.trace(rho)

    def full_coefficients_mps(self, tol_memory=20):
        """the state as a (d^L, 1) column vector, site 0 the slowest index (MPSClass.py:968-987, whose loop forgets to keep
        its reshape and fails for L > 2; this is what it is meant to return).  None when log2(d^L) - 5 > tol_memory."""
        if self.length * np.log2(self.phys_dim) - 5 > tol_memory:
            print('The memory cost of the total coefficients is too large; pass a larger tol_memory to calculate anyway')
            return None
        x = np.ones((1, 1))
        for t in self.mps:
            t = self._be.to_numpy(t) if hasattr(t, 'data_ptr') else np.asarray(t)
            x = x.dot(t.reshape(t.shape[0], -1)).reshape(-1, t.shape[2])
        return x.reshape(-1, 1)

    def effective_hamiltonian_dmrg(self, p, index1, index2, coeff1, coeff2, tol=1e-12):
        """dense H_eff of site p as an (n, n) numpy array, n = a*d*b (MPSClass.py:532-580; the eig_way = 0 object).  Built by
        applying the matvec plan to the columns of the identity, so it is the very operator the eigensolver sees; meant
        for checks at small n (refuses n > 4096)."""
        self._ensure_device()
        self.correct_orthogonal_center(p)
        plan = self.effective_hamiltonian_plan(p, index1, index2, coeff1, coeff2, tol=tol)
        shape = tuple(self.mps[p].shape)
        n = int(np.prod(shape))
        if n > 4096:
            plan.destroy()
            raise ValueError('effective_hamiltonian_dmrg: n = %d is too large for a dense matrix (limit 4096)' % n)
        h = np.zeros((n, n))
        e = np.zeros(n)
        for j in range(n):
            e[:] = 0.0
            e[j] = 1.0
            h[:, j] = self._be.to_numpy(plan.matvec(self._be.from_numpy(e.reshape(shape)), 0.0, 1.0)).reshape(-1)
        plan.destroy()
        return h

    # ---- checking functions (MPSClass.py:1024-1088) ----
    def check_orthogonality_by_tensors(self, tol=1e-12, is_print=True):
        """sites whose isometry property contradicts self.orthogonality (-1: left-, 1: right-orthonormal)"""
        bad = []
        for n in range(self.length):
            t = self._be.to_numpy(self.mps[n]) if hasattr(self.mps[n], 'data_ptr') else np.asarray(self.mps[n])
            o = int(np.ravel(self.orthogonality)[n])
            if o == -1:
                m = t.reshape(-1, t.shape[2])
                ok = np.abs(m.T @ m - np.eye(m.shape[1])).max() <= tol
            elif o == 1:
                m = t.reshape(t.shape[0], -1)
                ok = np.abs(m @ m.T - np.eye(m.shape[0])).max() <= tol
            else:
                ok = True
            if not ok:
                bad.append(n)
        if is_print:
            print('The orthogonality of all tensors are marked correctly by self.orthogonality' if not bad
                  else 'In self.orthogonality, the orthogonality of the following tensors is incorrect: ' + str(bad))
        return bad

    def check_virtual_bond_dimensions(self):
        bad = [n for n in range(1, self.length)
               if self.virtual_dim[n] != self.mps[n].shape[0] or self.virtual_dim[n] != self.mps[n - 1].shape[2]]
        for n in bad:
            print('BondDimError: inconsistent dimension detected for the %d-th virtual bond' % n)
        return bad

    def check_mps_norm1(self, if_print=False):
        norm = self.norm_mps()
        if abs(norm - 1) > 1e-14:
            print('The norm is MPS is %g away from 1' % abs(norm - 1))
        if if_print:
            print('The norm of MPS is %g' % norm)
        return norm

    # ---- housekeeping ----
    def print_general_info(self):
        print('DMRG & MPS (%s): B200-native drop-in of ranshiju/T-Nalg MpsOpenBoundaryClass; see README.md' % self.version)

    def report_yourself(self):
        print('center: ' + str(self.center))
        print('orthogonality:' + str(self.orthogonality.T))
        print('virtual bond dimensions: ' + str(self.virtual_dim))

    def clean_to_save(self):
        """drop caches and bring the tensors to the host as numpy arrays (MPSClass.py:1090-1098)"""
        on_dev = [i for i, t in enumerate(self.mps) if hasattr(t, 'data_ptr')]
        mps = [t if hasattr(t, 'data_ptr') else np.asarray(t) for t in self.mps]
        if on_dev:
            for i, h in zip(on_dev, self._be.to_numpy_many([mps[i] for i in on_dev])):
                mps[i] = h
        self.mps = mps
        self.effect_s = {'none': np.zeros(0)}
        self.effect_ss = {'none': np.zeros(0)}
        self.effective_id = {'none': np.zeros(0)}
        self.pos_effect_s = np.zeros((0, 3)).astype(int)
        self.pos_effect_ss = np.zeros((0, 5)).astype(int)
        self.pool = None
        self._env = None
        self._env_key = None


class MpsInfinite(MpsBasic):
    """White-style two-site iDMRG state on the CUDA kernels: the part of the reference's MpsInfinite
    (library/MPSClass.py:1560-1870) with dmrg_type='white', n_site=2, form='center_ort', symmetric environments.

    Same constructor signature and method names.  mps[1] is the central two-site tensor (chi, d, d, chi); mps[0] / mps[2] the
    left-orthonormal tensors of the last split; bath_op_onsite and effective_ops[n] the block Hamiltonian and the block
    operators of the growing half-chain.  update_central_tensor solves the two-site effective Hamiltonian
    (update_central_tensor_effective_ops_fh, :1688-1707 -- a term-summed two-site matvec, here one tn_effh_plan with the combined
    physical index d*d) with the device-resident Lanczos; update_ort_tensor_mps is the SVD truncation (:1676-1686) on the
    Jacobi kernel.  The MPO-form iDMRG (dmrg_type='mpo') is not provided (out of scope)."""

    def __init__(self, form, d, chi, D, n_tensor=3, n_site=1, spin='half', dmrg_type='mpo', way='qr', operators=None,
                 hamilt_index=None, is_symme_env=True, is_real=True, debug=False):
        MpsBasic.__init__(self)
        if dmrg_type != 'white' or n_site != 2 or form != 'center_ort':
            raise NotImplementedError("MpsInfinite: only dmrg_type='white', n_site=2, form='center_ort' run on the CUDA path")
        if hamilt_index is None:
            raise ValueError('MpsInfinite: hamilt_index (rows [op1, op2, coupling]) is required')
        self._be = _ops.backend()
        self.spin, self.n_tensor, self.n_site = spin, 3, n_site
        self.d, self.chi, self.D, self.form = d, chi, D, form
        self.decomp_way, self.is_symme_env, self.is_real, self._debug, self.dmrg_type = way, True, is_real, debug, dmrg_type
        self.orthogonality = np.array([-1, 0, 1])
        self.is_center_ort, self.is_canonical = True, False
        if operators is None:
            op_half = spin_operators(spin)
            self.operators = [op_half['id'], op_half['sx'], op_half['sy'], op_half['sz'], op_half['su'], op_half['sd']]
        else:
            self.operators = operators
        self.hamilt_index = np.asarray(hamilt_index, dtype=float).reshape(-1, 3)
        self.op_index = set()
        self.simplified_op_index()
        be = self._be
        # randomly initialised central tensor, same draw as initialize_imps (:1618-1624)
        psi = np.random.randn(chi, d, d, chi)
        psi /= np.linalg.norm(psi.reshape(-1))
        self.mps = [None, be.from_numpy(psi.reshape(chi, d * d, chi)), None]
        self.lm = [np.zeros(0)] * 3
        self.rho = None
        self.update_ort_tensor_mps('both')
        self.bath_op_onsite = be.from_numpy(np.eye(chi))
        self.effective_ops = [be.from_numpy(np.eye(chi)) for _ in range(len(self.operators))]
        self.stats = {'n_solves': 0, 'n_matvec': 0}
        self.lanczos_ncv, self.lanczos_max_restarts = 20, 2000

    def _real_op(self, n):
        o = np.asarray(self.operators[int(n)])
        if np.abs(np.imag(o)).max() != 0:
            raise ValueError('operator %d is complex; the real path supports real site operators only' % int(n))
        return np.real(o).astype(float)

    def simplified_op_index(self):
        """the operators whose block versions are needed (:1643-1649)"""
        self.op_index = {1, 3}
        for n in range(self.hamilt_index.shape[0]):
            self.op_index.add(int(self.hamilt_index[n, 0]))
            self.op_index.add(int(self.hamilt_index[n, 1]))

    def update_ort_tensor_mps(self, which='both', dc=None):
        """split the central tensor: theta (chi*d, d*chi) = U S Vh truncated to dc (:1676-1686);
        mps[0] = U (chi, d, dc), mps[2][b, s, k] = Vh[k, (s, b)]"""
        be = self._be
        chi, d = self.mps[1].shape[0], self.d
        b = self.mps[1].shape[2]
        dc = min(self.chi, chi * d) if dc is None else min(dc, chi * d)
        U, S, Vt = be.svd(self.mps[1].reshape(chi * d, d * b), k_keep=dc)
        self.mps[0] = U.contiguous().reshape(chi, d, dc)
        self.mps[2] = Vt.reshape(dc, d, b).permute(2, 1, 0).contiguous()
        self.lm[0] = be.to_numpy(S)

    def update_effective_ops(self, which='op_index'):
        """effective_ops[n] = block operator n grown by one site (:1762-1779), all of them in one batched tn_env_update"""
        idx = sorted(self.op_index) if which == 'op_index' else ([which] if isinstance(which, int) else list(which))
        idx = [n for n in idx if np.abs(np.imag(np.asarray(self.operators[n]))).max() == 0]
        outs = self._be.env_update(0, self.mps[0], [[(None, self._real_op(n))] for n in idx])
        for n, mat in zip(idx, outs):
            self.effective_ops[n] = mat

    def update_bath_onsite(self):
        """block Hamiltonian of the half-chain grown by one site (:1781-1790): one output with 1 + #terms links"""
        links = [(self.bath_op_onsite, None)]
        for n in range(self.hamilt_index.shape[0]):
            i1, i2, j = int(self.hamilt_index[n, 0]), int(self.hamilt_index[n, 1]), float(self.hamilt_index[n, 2])
            if j != 0.0:
                links.append((self.effective_ops[i1], j * self._real_op(i2)))
        op = self._be.env_update(0, self.mps[0], [links])[0]
        self.bath_op_onsite = (op + op.t()) / 2

    def effective_plan(self):
        """tn_effh_plan of the two-site effective Hamiltonian (update_central_tensor_effective_ops_fh, :1688-1707):
        bath (x) 1 + 1 (x) bath + sum_n j_n op1 (x) op2 on the pair + sum_n j_n [block(op1) (x) op2 (x) 1 + 1 (x) op2 (x) block(op1)]"""
        be, d, chi = self._be, self.d, self.mps[1].shape[0]
        eye = np.eye(d)
        M = np.zeros((d * d, d * d))
        ls, rs = {}, {}
        for n in range(self.hamilt_index.shape[0]):
            i1, i2, j = int(self.hamilt_index[n, 0]), int(self.hamilt_index[n, 1]), float(self.hamilt_index[n, 2])
            if j == 0.0:
                continue
            M += j * np.kron(self._real_op(i1), self._real_op(i2))
            ls.setdefault(i2, []).append((j, i1))
            rs.setdefault(i2, []).append((j, i1))
        LS, ls_ops, RS, rs_ops = [], [], [], []
        for i2, pairs in sorted(ls.items()):
            mat = be.lincomb([self.effective_ops[i1] for _, i1 in pairs], [c for c, _ in pairs])
            LS.append(mat)
            ls_ops.append(np.kron(self._real_op(i2), eye))
            RS.append(mat)
            rs_ops.append(np.kron(eye, self._real_op(i2)))
        return be.effh_plan((chi, d * d, self.mps[1].shape[2]), self.bath_op_onsite, self.bath_op_onsite, M, LS, ls_ops, RS, rs_ops)

    def update_central_tensor_effective_ops_fh(self, psi, tau):
        """psi -> psi - tau * H_eff psi on a (chi, d, d, chi) tensor given as numpy or device tensor (:1688-1707)"""
        be = self._be
        dev = hasattr(psi, 'data_ptr')
        x = psi if dev else be.from_numpy(np.real(np.asarray(psi)))
        plan = self.effective_plan()
        y = plan.matvec(x.reshape(plan.shape), 1.0, -float(tau)).clone()
        plan.destroy()
        return y if dev else be.to_numpy(y).reshape(np.asarray(psi).shape)

    def effective_hamilt_from_op(self):
        """dense two-site effective Hamiltonian (:1709-1729), through the plan (small chi only)"""
        be = self._be
        plan = self.effective_plan()
        n = int(np.prod(plan.shape))
        if n > 4096:
            plan.destroy()
            raise ValueError('effective_hamilt_from_op: dimension %d too large for a dense matrix' % n)
        h = np.zeros((n, n))
        for j in range(n):
            e = np.zeros(n)
            e[j] = 1.0
            h[:, j] = be.to_numpy(plan.matvec(be.from_numpy(e.reshape(plan.shape)), 0.0, 1.0)).reshape(-1)
        plan.destroy()
        return h

    def update_central_tensor(self, inputs):
        """dominant eigenvector of 1 - tau * H_eff as the new central tensor (:1840-1850); inputs = (tau, way) -- both ways
        use the operator (plan) form here, the dense 'full' matrix of the reference is the same operator"""
        tau = inputs[0] if isinstance(inputs, (tuple, list)) else float(inputs)
        plan = self.effective_plan()
        # The previous central tensor has a definite mirror parity; when the ground state of the grown block sits in the other
        # sector it is EXACTLY orthogonal to it and no Krylov method started from it can reach it (the reference's ARPACK call
        # gets there through round-off).  A 1e-6 admixture of a fixed pseudo-random vector keeps every sector in the Krylov space.
        v0 = self.mps[1].reshape(-1)
        noise = np.random.RandomState(1000 + self.stats['n_solves']).randn(v0.numel())
        v0 = v0 + self._be.from_numpy(noise * (1e-6 / np.sqrt(noise.size)))
        lam, vec, n_mv, resid, ok = self._be.lanczos(plan, tau, v0, 1e-14, ncv=self.lanczos_ncv,
                                                     max_restarts=self.lanczos_max_restarts)
        plan.destroy()
        self.stats['n_solves'] += 1
        self.stats['n_matvec'] += n_mv
        self.mps[1] = vec.reshape(self.mps[1].shape)

    def rho_from_central_tensor(self):
        """two-site reduced density matrix of the central pair (:1852-1868), (d*d, d*d) on the host"""
        t = self._be.to_numpy(self.mps[1])          # (chi, d*d, chi)
        self.rho = np.einsum('asb,atb->st', t, t)
        return self.rho

    def observe_energy(self, h):
        return float(np.trace(self.rho.dot(np.real(h))))

    def check_orthogonality_mps(self):
        a = self._be.to_numpy(self.mps[0])
        b = self._be.to_numpy(self.mps[2])
        k = a.shape[2]
        return (np.abs(np.einsum('asx,asy->xy', a, a) - np.eye(k)).max() < 1e-12 and
                np.abs(np.einsum('asx,asy->xy', b, b) - np.eye(k)).max() < 1e-12)


class MpsStandardTEBD(MpsOpenBoundaryClass):
    """imaginary-time TEBD with two-body gates on the CUDA kernels (library/MPSClass.py:1408-1447): evolve_gate_tebd applies the
    two halves of an SVD-split gate to sites p1 < p2 (growing the bonds in between by the gate rank), truncate_mps_tebd brings
    the bonds back to chi with SVD-truncating gauge moves (tn_svd_jacobi with k_keep)."""

    def __init__(self, length, d, chi, spin='half', ini_way='r', evolve_way='gate'):
        MpsOpenBoundaryClass.__init__(self, length, d, chi, spin=spin, way='qr', ini_way=ini_way, operators=None, debug=False,
                                      is_parallel=False, is_save_op=False, eig_way=0, par_pool=None, is_env_parallel_lmr=False)
        self.chi = chi

    def evolve_tensor(self, p, gate, enlarge_which_bond):
        """mps[p] <- gate (d, dd, d) applied on the physical bond, the gate's middle index merged into the right (2) or left (0)
        virtual bond (library/MPSClass.py:894-907)"""
        import torch
        self._ensure_device()
        be = self._be
        gate = np.real(np.asarray(gate))
        d, dd = gate.shape[:2]
        T = self.mps[p]
        chi1, _, chi2 = T.shape
        parts = [be.site_op(T, gate[:, k, :]) for k in range(dd)]      # parts[k][a, s, b] = sum_s' gate[s, k, s'] T[a, s', b]
        if enlarge_which_bond == 2:
            self.mps[p] = torch.stack(parts, dim=2).reshape(chi1, d, dd * chi2).contiguous()     # (a, s, k, b)
            self.virtual_dim[p + 1] = dd * chi2
        else:
            self.mps[p] = torch.stack(parts, dim=0).reshape(dd * chi1, d, chi2).contiguous()     # (k, a, s, b)
            self.virtual_dim[p] = dd * chi1
        self.orthogonality[p] = 0

    def evolve_gate_tebd(self, p1, p2, gates):
        import torch
        if p1 > p2:
            p1, p2 = p2, p1
        dd = np.asarray(gates[0]).shape[1]
        self.evolve_tensor(p1, gates[0], 2)
        for n in range(p1 + 1, p2):   # identity string of dimension dd through the sites in between (:1421-1427)
            t = self.mps[n]
            a, d, b = t.shape
            eye = torch.eye(dd, dtype=t.dtype, device=t.device)
            self.mps[n] = torch.einsum('kl,asb->kaslb', eye, t).reshape(dd * a, d, dd * b).contiguous()
            self.virtual_dim[n + 1] = dd * b
            self.orthogonality[n] = 0
        self.evolve_tensor(p2, gates[1], 0)

    def truncate_mps_tebd(self, p1, p2):
        if p1 > p2:
            p1, p2 = p2, p1
        c0 = self.center
        if c0 < p1:
            self.orthogonalize_mps(p2, p1, normalize=True, is_trun=False)
            self.orthogonalize_mps(c0, p1, normalize=True, is_trun=False)
        elif c0 > p2:
            self.orthogonalize_mps(self.center, p1, normalize=True, is_trun=False)
        else:
            self.orthogonalize_mps(p2, p1, normalize=True, is_trun=False)
        self.orthogonalize_mps(p1, p2, normalize=True, is_trun=True, chi=self.chi)
        self.center = p2

    def norm_mps(self, if_normalize=False):
        norm = MpsOpenBoundaryClass.norm_mps(self)
        if if_normalize and self.center > -0.1:
            self.mps[self.center] = self.mps[self.center] / norm
        return norm

    def observe_bond_energy_from_jxyz(self, pos2, jx, jy, jz, tol=1e-20):
        """eb[n] = jx <sx sx> + jy <sy sy> + jz <sz sz> on the pairs pos2 (library/MPSClass.py:1184-1194).  sy is complex; its
        correlator is measured through the real ladder operators, sy (x) sy = -(su - sd) (x) (su - sd) / 4."""
        pos2 = np.asarray(pos2, dtype=int)
        terms, weight = [], []
        for n in range(pos2.shape[0]):
            a, b = sorted((int(pos2[n, 0]), int(pos2[n, 1])))
            for o1, o2, c in ((1, 1, jx), (3, 3, jz), (4, 4, -jy / 4), (4, 5, jy / 4), (5, 4, jy / 4), (5, 5, -jy / 4)):
                if abs(c) > tol:
                    terms.append(((a, o1), (b, o2)))
                    weight.append((n, c))
        vals = self._expect(terms)
        eb = np.zeros((pos2.shape[0], 1))
        for (n, c), v in zip(weight, vals):
            eb[n] += c * v
        return eb
