"""Parity of every C-ABI entry point against the numpy oracle on identical seeded inputs (run on the B200 box)."""
import ctypes as C

import numpy as np
import pytest

from oracle import dmrg_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def be():
    from tnalg_b200 import ops
    return ops.backend()


def rel_err(x, ref):
    return np.abs(np.asarray(x) - ref).max() / max(np.abs(ref).max(), 1e-300)


def chain_gemm(be, mode, M, N, K, d, problems, deterministic):
    """problems: list of (alpha, accumulate, C0 (M,N) numpy, [ (A, B, op or None) ]) -> list of results"""
    import torch
    from tnalg_b200 import _lib as L
    n_links = sum(len(p[3]) for p in problems)
    probs = (L.TnProblem * len(problems))()
    links = (L.TnLink * n_links)()
    keep, outs, li = [], [], 0
    for q, (alpha, acc, C0, lk) in enumerate(problems):
        Cd = be.from_numpy(C0)
        outs.append(Cd)
        probs[q].C, probs[q].alpha, probs[q].link_begin, probs[q].link_count, probs[q].accumulate = Cd.data_ptr(), alpha, li, len(lk), acc
        for A, B, op in lk:
            Ad, Bd = be.from_numpy(A), be.from_numpy(B)
            keep += [Ad, Bd]
            links[li].A, links[li].B, links[li].has_op = Ad.data_ptr(), Bd.data_ptr(), 0 if op is None else 1
            if op is not None:
                for i, v in enumerate(np.asarray(op).reshape(-1)):
                    links[li].op[i] = v
            li += 1
    lda = M if mode == 2 else K
    ldb = K if mode == 1 else N
    ws = torch.empty(be.lib.tn_chain_gemm_workspace_bytes(len(problems), n_links) + 256, dtype=torch.uint8, device=be.device)
    L.check(be.lib.tn_chain_gemm(mode, M, N, K, d, lda, ldb, N, probs, len(problems), links, n_links, deterministic,
                                 C.c_void_p(ws.data_ptr()), ws.numel(), be.stream()))
    torch.cuda.synchronize()
    return [be.to_numpy(o) for o in outs]


def ref_link(mode, d, A, B, op):
    if mode == 0:      # A (M,K) . opcol(B (K, d*Ny))
        if op is not None:
            K, N = B.shape
            B = np.einsum('st,kty->ksy', op, B.reshape(K, d, N // d)).reshape(K, N)
        return A @ B
    if mode == 1:      # oprow(A (Mx*d, K)) . B(N,K)^T
        if op is not None:
            M, K = A.shape
            A = np.einsum('st,xtk->xsk', op, A.reshape(M // d, d, K)).reshape(M, K)
        return A @ B.T
    return A.T @ B     # A (K,M), B (K,N)


SHAPES = [(1, 2, 1), (2, 4, 2), (5, 6, 3), (16, 16, 16), (33, 18, 7), (64, 64, 64), (100, 36, 50), (130, 258, 70),
          (256, 256, 96), (257, 129, 33)]


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('d', [1, 2, 3, 4])
@pytest.mark.parametrize('deterministic', [1, 0])
def test_chain_gemm_shapes(be, mode, d, deterministic):
    rng = np.random.RandomState(100 * mode + 10 * d + deterministic)
    if mode == 2 and d > 1:
        pytest.skip('TN has no operator')
    for (M, N, K) in SHAPES:
        if mode == 0:
            N = N * d
        if mode == 1:
            M = M * d
        shapeA = (K, M) if mode == 2 else (M, K)
        shapeB = (N, K) if mode == 1 else (K, N)
        problems, refs = [], []
        for q, n_links in enumerate([3, 1]):
            lk = []
            ref = np.zeros((M, N))
            for l in range(n_links):
                A, B = rng.randn(*shapeA), rng.randn(*shapeB)
                op = rng.randn(d, d) if (d > 1 and l != 1) else None
                lk.append((A, B, op))
                ref += ref_link(mode, d, A, B, op)
            alpha, acc = (0.7, 1) if q == 0 else (-1.3, 0)
            C0 = rng.randn(M, N)
            refs.append(alpha * ref + (C0 if acc else 0))
            problems.append((alpha, acc, C0, lk))
        outs = chain_gemm(be, mode, M, N, K, d, problems, deterministic)
        for o, r in zip(outs, refs):
            assert rel_err(o, r) < 1e-13, (mode, d, M, N, K)


def test_chain_gemm_large_split_and_deterministic(be):
    """chi=512-like sizes: the 128x128 configuration, stream-K split vs whole tiles, run-to-run determinism"""
    rng = np.random.RandomState(7)
    M, N, K, d = 512, 512, 512, 2
    lk = [(rng.randn(M, K), rng.randn(N, K), rng.randn(d, d) if l % 2 else None) for l in range(4)]
    ref = sum(ref_link(1, d, *x) for x in lk)
    C0 = np.zeros((M, N))
    o_split = chain_gemm(be, 1, M, N, K, d, [(1.0, 0, C0, lk)], 0)[0]
    o_det1 = chain_gemm(be, 1, M, N, K, d, [(1.0, 0, C0, lk)], 1)[0]
    o_det2 = chain_gemm(be, 1, M, N, K, d, [(1.0, 0, C0, lk)], 1)[0]
    assert rel_err(o_split, ref) < 1e-13 and rel_err(o_det1, ref) < 1e-13
    assert np.array_equal(o_det1, o_det2)


@pytest.mark.parametrize('shape', [(1, 2, 2), (2, 2, 4), (4, 2, 8), (16, 2, 16), (7, 3, 9), (64, 2, 32), (96, 2, 96), (130, 2, 128), (20, 4, 12)])
def test_env_update_vs_oracle(be, shape):
    rng = np.random.RandomState(sum(shape))
    a, d, b = shape
    T = rng.randn(a, d, b)
    ops = [rng.randn(d, d) for _ in range(3)]
    for direction, e_dim, fn in ((0, a, orc.transfer_l2r), (1, b, orc.transfer_r2l)):
        Es = [rng.randn(e_dim, e_dim) for _ in range(3)]
        outputs_np = [
            [(Es[0], None), (Es[1], ops[0]), (None, ops[1])],   # an HL-like chain
            [(Es[2], None)],                                    # plain transfer
            [(None, ops[2])],                                   # new operator, identity environment
            [(None, None)],                                     # identity transfer (density chain)
            [(Es[1], ops[2])],
        ]
        outputs = [[(None if E is None else be.from_numpy(E), op) for E, op in links] for links in outputs_np]
        res = be.env_update(direction, be.from_numpy(T), outputs)
        for r, links in zip(res, outputs_np):
            ref = sum(fn(T, op, E) for E, op in links)
            assert rel_err(be.to_numpy(r), ref) < 1e-13, (shape, direction)


def random_groups(rng, a, d, b, n_ls, n_rs, n_x, with_h=True, with_m=True):
    sym = lambda n: (lambda g: (g + g.T) / 2)(rng.randn(n, n))  # noqa: E731
    return {'HL': sym(a) if with_h else None, 'HR': sym(b) if with_h else None, 'M': rng.randn(d, d) if with_m else None,
            'LS': [rng.randn(a, a) for _ in range(n_ls)], 'ls_ops': [rng.randn(d, d) for _ in range(n_ls)],
            'RS': [rng.randn(b, b) for _ in range(n_rs)], 'rs_ops': [rng.randn(d, d) for _ in range(n_rs)],
            'XL': [rng.randn(a, a) for _ in range(n_x)], 'XR': [rng.randn(b, b) for _ in range(n_x)],
            'x_coeff': list(rng.randn(n_x))}


def dense_from_groups(g, shape):
    """H_eff = sum kron(kron(L, M), R) (effective_hamiltonian_dmrg, MPSClass.py:532-580)"""
    a, d, b = shape
    Ia, Id, Ib = np.eye(a), np.eye(d), np.eye(b)
    k3 = lambda x, y, z: np.kron(np.kron(x, y), z)  # noqa: E731
    H = np.zeros((a * d * b, a * d * b))
    if g['HL'] is not None:
        H += k3(g['HL'], Id, Ib)
    if g['HR'] is not None:
        H += k3(Ia, Id, g['HR'])
    if g['M'] is not None:
        H += k3(Ia, g['M'], Ib)
    for E, op in zip(g['LS'], g['ls_ops']):
        H += k3(E, op, Ib)
    for E, op in zip(g['RS'], g['rs_ops']):
        H += k3(Ia, op, E)
    for c, l, r in zip(g['x_coeff'], g['XL'], g['XR']):
        H += c * k3(l, Id, r)
    return H


def gpu_plan(be, shape, g, rank=0, world=1):
    t = lambda x: None if x is None else be.from_numpy(x)  # noqa: E731
    return be.effh_plan(shape, t(g['HL']), t(g['HR']), g['M'], [t(x) for x in g['LS']], g['ls_ops'], [t(x) for x in g['RS']],
                        g['rs_ops'], [t(x) for x in g['XL']], [t(x) for x in g['XR']], g['x_coeff'], rank=rank, world=world)


@pytest.mark.parametrize('shape,counts', [((1, 2, 2), (0, 3, 0)), ((2, 2, 1), (3, 0, 0)), ((4, 2, 8), (3, 3, 2)), ((16, 2, 16), (3, 3, 0)),
                                          ((9, 3, 27), (2, 3, 4)), ((64, 2, 64), (3, 3, 9)), ((128, 2, 128), (3, 3, 5)),
                                          ((256, 2, 256), (3, 3, 2)), ((100, 2, 60), (1, 2, 3)), ((33, 4, 47), (4, 3, 5)), ((8, 4, 8), (2, 2, 1))])
def test_matvec_vs_oracle(be, shape, counts):
    from tests.cpu_backend import CpuPlan
    rng = np.random.RandomState(sum(shape) + sum(counts))
    a, d, b = shape
    g = random_groups(rng, a, d, b, *counts)
    ref_plan = CpuPlan(shape, g)
    plan = gpu_plan(be, shape, g)
    x = rng.randn(a, d, b)
    hx = ref_plan.apply(x.reshape(-1))
    y = be.to_numpy(plan.matvec(be.from_numpy(x), 0.0, 1.0)).reshape(-1)
    assert rel_err(y, hx) < 1e-13, shape
    tau = 1e-4
    y = be.to_numpy(plan.matvec(be.from_numpy(x), 1.0, -tau)).reshape(-1)
    assert np.abs(y - (x.reshape(-1) - tau * hx)).max() < 1e-15 * max(1.0, np.abs(hx).max())
    assert plan.flops_algorithmic == ref_plan.flops_algorithmic
    # sharded plans sum to the full operator (multi-GPU term sharding)
    parts = [be.to_numpy(gpu_plan(be, shape, g, rank=r, world=3).matvec(be.from_numpy(x), 1.0, -tau)).reshape(-1) for r in range(3)]
    assert np.abs(sum(parts) - y).max() < 1e-14 * max(1.0, np.abs(y).max())


@pytest.mark.parametrize('shape,counts', [((384, 2, 320), (3, 3, 10)), ((512, 2, 512), (3, 3, 4)), ((300, 2, 260), (2, 3, 9)),
                                          ((258, 3, 264), (3, 2, 6)), ((640, 2, 192), (3, 3, 0)), ((200, 4, 168), (4, 4, 7))])
@pytest.mark.parametrize('tma', [True, False])
def test_matvec_large_config_vs_oracle(be, shape, counts, tma, monkeypatch):
    """shapes that are scheduled on the 128x64 configuration: the TMA-staged kernel (default) and the cp.async kernel
    (TNALG_NO_TMA=1) against the numpy oracle, including ragged edges (zero fill) and d = 3"""
    from tests.cpu_backend import CpuPlan
    if not tma:
        monkeypatch.setenv('TNALG_NO_TMA', '1')
    else:
        monkeypatch.delenv('TNALG_NO_TMA', raising=False)
    rng = np.random.RandomState(sum(shape) + sum(counts))
    a, d, b = shape
    g = random_groups(rng, a, d, b, *counts)
    x = rng.randn(a, d, b)
    ref = CpuPlan(shape, g).apply(x.reshape(-1))
    plan = gpu_plan(be, shape, g)
    # both stages are TMA-staged when both are scheduled on the 128x64 configuration; the last shape (no crossing terms,
    # narrow right bond) legitimately runs its right stage on the small cp.async configuration
    expect = 0 if not tma else (1 if counts[2] == 0 else 3)
    assert plan.uses_tma == expect, 'expected kernel mask %d, got %d' % (expect, plan.uses_tma)
    y = be.to_numpy(plan.matvec(be.from_numpy(x), 0.0, 1.0)).reshape(-1)
    assert rel_err(y, ref) < 1e-13, (shape, tma)
    y2 = be.to_numpy(plan.matvec(be.from_numpy(x), 1.0, -0.25)).reshape(-1)
    assert rel_err(y2, x.reshape(-1) - 0.25 * ref) < 1e-13
    parts = [be.to_numpy(gpu_plan(be, shape, g, rank=r, world=2).matvec(be.from_numpy(x), 0.0, 1.0)).reshape(-1) for r in range(2)]
    assert rel_err(sum(parts), ref) < 1e-13


def test_matvec_without_any_block(be):
    """a site with only an on-site field (no environment blocks at all) still works"""
    rng = np.random.RandomState(3)
    g = random_groups(rng, 4, 2, 4, 0, 0, 0, with_h=False)
    x = rng.randn(4, 2, 4)
    y = be.to_numpy(gpu_plan(be, (4, 2, 4), g).matvec(be.from_numpy(x), 0.0, 1.0))
    assert rel_err(y, np.einsum('st,atb->asb', g['M'], x)) < 1e-14


@pytest.mark.parametrize('shape,counts', [((1, 2, 2), (0, 3, 0)), ((2, 2, 4), (3, 3, 1)), ((8, 2, 8), (3, 3, 2)), ((24, 2, 24), (3, 3, 4)),
                                          ((40, 2, 40), (3, 3, 3))])
@pytest.mark.parametrize('tol', [1e-5, 1e-12])
@pytest.mark.parametrize('fused', [True, False])
def test_lanczos_vs_dense_eigh(be, shape, counts, tol, fused, monkeypatch):
    """dominant eigenpair of 1 - tau*H_eff: eigenvalue, residual and (for tight tol) the eigenvector itself; both the fused
    cooperative re-orthogonalisation kernel (small vectors) and the streaming kernels with the DGKS-conditional second pass"""
    from tests.cpu_backend import CpuPlan
    if fused:
        monkeypatch.delenv('TNALG_NO_FUSED_ORTH', raising=False)
    else:
        monkeypatch.setenv('TNALG_NO_FUSED_ORTH', '1')
    rng = np.random.RandomState(11 + sum(shape))
    a, d, b = shape
    g = random_groups(rng, a, d, b, *counts)
    # make every piece symmetric so that H_eff is symmetric, as it is in DMRG
    for k in ('LS', 'RS', 'XL', 'XR'):
        g[k] = [(m + m.T) / 2 for m in g[k]]
    for k in ('ls_ops', 'rs_ops'):
        g[k] = [(m + m.T) / 2 for m in g[k]]
    g['M'] = (g['M'] + g['M'].T) / 2
    n = a * d * b
    H = dense_from_groups(g, shape)
    assert np.abs(H - H.T).max() < 1e-12
    xchk = rng.randn(n)
    assert np.abs(H @ xchk - CpuPlan(shape, g).apply(xchk)).max() < 1e-10
    tau = 1e-4
    w, v = np.linalg.eigh(np.eye(n) - tau * H)
    k = int(np.argmax(np.abs(w)))
    plan = gpu_plan(be, shape, g)
    v0 = rng.randn(n)
    lam, vec, n_mv, resid, ok = be.lanczos(plan, tau, be.from_numpy(v0), tol, ncv=20, max_restarts=500)
    vec = be.to_numpy(vec)
    assert ok
    assert abs(np.linalg.norm(vec) - 1) < 1e-13
    r = (np.eye(n) - tau * H) @ vec - lam * vec
    assert np.linalg.norm(r) <= 2 * tol * abs(lam) + 1e-14
    if tol < 1e-10:
        assert abs(lam - w[k]) < 1e-13
        assert min(np.linalg.norm(vec - v[:, k]), np.linalg.norm(vec + v[:, k])) < 1e-6
    assert n_mv >= min(n, 2)


@pytest.mark.parametrize('shape', [(1, 1), (2, 1), (1, 5), (4, 4), (8, 3), (3, 8), (37, 21), (64, 64), (130, 64), (64, 130), (256, 128),
                                   (33, 33), (100, 70), (512, 512), (700, 260)])
def test_svd_jacobi_vs_lapack(be, shape):
    rng = np.random.RandomState(shape[0] * 7 + shape[1])
    m, n = shape
    # graded singular values: Schmidt spectra span many decades
    k = min(m, n)
    A = np.linalg.qr(rng.randn(m, k))[0] @ np.diag(np.logspace(0, -12, k)) @ np.linalg.qr(rng.randn(n, k))[0].T
    U, S, Vt = [be.to_numpy(x) for x in be.svd(be.from_numpy(A))]
    s_ref = np.linalg.svd(A, compute_uv=False)
    assert np.abs(S - s_ref).max() <= 1e-10 * s_ref.max()          # north_star tolerance
    big = s_ref > 1e-8 * s_ref.max()                                # A itself is only known to 1e-16 absolute
    assert np.abs(S[big] / s_ref[big] - 1).max() < 1e-7             # Jacobi keeps relative accuracy of small values
    assert np.abs(U.T @ U - np.eye(k)).max() < 1e-12
    assert np.abs(Vt @ Vt.T - np.eye(k)).max() < 1e-12
    assert np.abs((U * S) @ Vt - A).max() < 1e-13
    if k >= 4:
        kk = k // 2
        U2, S2, Vt2 = [be.to_numpy(x) for x in be.svd(be.from_numpy(A), k_keep=kk)]
        assert U2.shape == (m, kk) and Vt2.shape == (kk, n)
        assert np.abs(S2 - s_ref[:kk]).max() <= 1e-10 * s_ref.max()
        assert np.abs((U2 * S2) @ Vt2 - (U[:, :kk] * S[:kk]) @ Vt[:kk]).max() < 1e-12


def test_svd_ill_conditioned_preconditioned(be):
    """generic matrix with singular values over ten decades: QR-preconditioned Jacobi converges in a few sweeps and keeps
    every singular value to high relative accuracy (the raw kernel needs many more sweeps)"""
    rng = np.random.RandomState(2)
    m, n = 384, 192
    s = np.logspace(0, -10, n)
    A = (np.linalg.qr(rng.randn(m, n))[0] * s) @ np.linalg.qr(rng.randn(n, n))[0].T
    U, S, Vt = [be.to_numpy(x) for x in be.svd(be.from_numpy(A))]
    assert be.last_svd_sweeps <= 14
    assert np.abs(S / s - 1).max() < 1e-5          # A itself carries ~1e-16 absolute noise: 1e-6 relative at 1e-10
    assert np.abs(S[:n // 2] / s[:n // 2] - 1).max() < 1e-10
    assert np.abs(U.T @ U - np.eye(n)).max() < 1e-12 and np.abs(Vt @ Vt.T - np.eye(n)).max() < 1e-12
    assert np.abs((U * S) @ Vt - A).max() < 1e-13
    U0, S0, Vt0 = [be.to_numpy(x) for x in be.svd(be.from_numpy(A), precondition=False)]
    assert np.abs(S0 - s).max() < 1e-10 and be.last_svd_sweeps > 14


@pytest.mark.parametrize('shape', [(5, 3), (64, 64), (130, 64), (300, 300)])
def test_svd_unblocked_kernel_agrees(be, shape, monkeypatch):
    """the unblocked kernel (rows in global memory; used when two rows do not fit shared memory) gives the same SVD"""
    rng = np.random.RandomState(shape[0] + shape[1])
    A = rng.randn(*shape)
    U, S, Vt = [be.to_numpy(x) for x in be.svd(be.from_numpy(A))]
    monkeypatch.setenv('TNALG_SVD_UNBLOCKED', '1')
    U0, S0, Vt0 = [be.to_numpy(x) for x in be.svd(be.from_numpy(A))]
    s_ref = np.linalg.svd(A, compute_uv=False)
    assert np.abs(S - s_ref).max() < 1e-13 * s_ref.max() and np.abs(S0 - s_ref).max() < 1e-13 * s_ref.max()
    assert np.abs((U * S) @ Vt - A).max() < 1e-12 and np.abs((U0 * S0) @ Vt0 - A).max() < 1e-12


def test_svd_rank_deficient_two_site_wavefunction(be):
    """theta of a two-site update: numerical rank chi out of 2*chi, the rest is round-off -- kept triplets exact, all
    vectors orthonormal, convergence within the sweep budget"""
    rng = np.random.RandomState(12)
    chi = 96
    s = np.concatenate([np.logspace(0, -9, chi), 1e-17 * rng.rand(chi)])
    A = (np.linalg.qr(rng.randn(2 * chi, 2 * chi))[0] * s) @ np.linalg.qr(rng.randn(2 * chi, 2 * chi))[0].T
    U, S, Vt = [be.to_numpy(x) for x in be.svd(be.from_numpy(A), k_keep=chi)]
    assert be.last_svd_sweeps <= 20
    assert np.abs(S[:chi // 2] / s[:chi // 2] - 1).max() < 1e-9 and np.abs(S - s[:chi]).max() < 1e-14
    assert np.abs(U.T @ U - np.eye(chi)).max() < 1e-12 and np.abs(Vt @ Vt.T - np.eye(chi)).max() < 1e-12
    assert np.abs((U * S) @ Vt - A).max() < 1e-13


@pytest.mark.parametrize('n', [256, 512])
def test_svd_spectrum_decaying_through_roundoff_converges(be, n):
    """two-site wavefunction late in a sweep: singular values decay smoothly from 1 to 1e-22, a third of them below the
    round-off level of the matrix; the truncating call must converge (regression: a shortcut that skipped pairs of
    negligible rows stalled the blocked kernels on such inputs) and keep the leading values exact"""
    rng = np.random.RandomState(n)
    s = np.logspace(0, -22, n)
    A = (np.linalg.qr(rng.randn(n, n))[0] * s) @ np.linalg.qr(rng.randn(n, n))[0].T
    U, S, Vt = [be.to_numpy(x) for x in be.svd(be.from_numpy(A), k_keep=n // 2)]
    assert be.last_svd_sweeps <= 40
    lead = s > 1e-9
    assert np.abs(S[lead[:n // 2]] / s[:n // 2][lead[:n // 2]] - 1).max() < 1e-6
    assert np.abs(U.T @ U - np.eye(n // 2)).max() < 1e-12 and np.abs(Vt @ Vt.T - np.eye(n // 2)).max() < 1e-12


@pytest.mark.parametrize('k_keep', [30, 60, 96])
def test_svd_truncation_beyond_numerical_rank_stays_orthonormal(be, k_keep):
    """rank 40 matrix (the rest is exactly at round-off level): asking for fewer, more, or all triplets must give orthonormal
    vectors -- pairs of negligible rows are only skipped when enough rows stay above the round-off level"""
    rng = np.random.RandomState(21)
    n, r = 96, 40
    A = (np.linalg.qr(rng.randn(n, r))[0] * np.logspace(0, -6, r)) @ np.linalg.qr(rng.randn(n, r))[0].T
    for precondition in (True, False):
        U, S, Vt = [be.to_numpy(x) for x in be.svd(be.from_numpy(A), k_keep=k_keep, precondition=precondition)]
        assert U.shape == (n, k_keep) and Vt.shape == (k_keep, n)
        kk = min(k_keep, r)
        assert np.abs(S[:kk] / np.logspace(0, -6, r)[:kk] - 1).max() < 1e-9
        assert np.abs(U[:, :kk].T @ U[:, :kk] - np.eye(kk)).max() < 1e-12
        assert np.abs(Vt[:kk] @ Vt[:kk].T - np.eye(kk)).max() < 1e-12
        assert np.abs(Vt @ Vt.T - np.eye(k_keep)).max() < 1e-10     # accumulated rotations: orthonormal for every row
        u0, s0, v0 = np.linalg.svd(A)
        assert np.abs((U[:, :kk] * S[:kk]) @ Vt[:kk] - (u0[:, :kk] * s0[:kk]) @ v0[:kk]).max() < 1e-13


def test_svd_matches_cusolver_baseline(be):
    """cuSOLVER (torch.linalg.svd) is kept only as a checked baseline (north_star)"""
    import torch
    rng = np.random.RandomState(5)
    A = be.from_numpy(rng.randn(192, 96))
    U, S, Vt = be.svd(A)
    S_ref = torch.linalg.svdvals(A)
    assert float((S - S_ref).abs().max()) <= 1e-12 * float(S_ref.max())


def test_two_site_svd_truncate_vs_oracle(be):
    """a12: two-site (a*d x d*b) SVD truncated to chi (library/MPSClass.py:1676-1686)"""
    from tnalg_b200 import TensorBasicModule as T
    rng = np.random.RandomState(9)
    theta = rng.randn(12, 2, 2, 10)
    U, lm, Vh = T.svd_truncate_two_site(theta, 8)
    U0, lm0, Vh0 = orc.svd_truncate_two_site(theta, 8)
    assert U.shape == (12, 2, 8) and Vh.shape == (8, 2, 10)
    assert np.abs(lm - lm0).max() <= 1e-10 * lm0.max()
    # subspaces agree up to sign
    assert np.abs(np.abs(np.einsum('asb,asc->bc', U, U0)) - np.eye(8)).max() < 1e-9
    assert np.abs(np.einsum('ayb,b,bxc->ayxc', U, lm, Vh) - np.einsum('ayb,b,bxc->ayxc', U0, lm0, Vh0)).max() < 1e-10


def test_small_ops(be):
    rng = np.random.RandomState(1)
    xs = [rng.randn(50, 50) for _ in range(20)]
    cs = rng.randn(20)
    out = be.to_numpy(be.lincomb([be.from_numpy(x) for x in xs], cs))
    assert rel_err(out, sum(c * x for c, x in zip(cs, xs))) < 1e-14
    T, op = rng.randn(9, 3, 11), rng.randn(3, 3)
    assert rel_err(be.to_numpy(be.site_op(be.from_numpy(T), op)), np.einsum('st,atb->asb', op, T)) < 1e-14
    x, y = rng.randn(100003), rng.randn(100003)
    E = rng.randn(77, 77)
    vals = be.scalars_to_host([be.dot(be.from_numpy(x), be.from_numpy(y)), be.trace(be.from_numpy(E))])
    assert abs(vals[0] - x @ y) < 1e-10 and abs(vals[1] - np.trace(E)) < 1e-12
    for bond, mat in ((0, rng.randn(9, 5)), (2, rng.randn(11, 4)), (1, rng.randn(3, 3))):
        got = be.to_numpy(be.mode_product(be.from_numpy(T), be.from_numpy(mat), bond))
        assert rel_err(got, orc.mode_product(T, mat, bond)) < 1e-13, bond
