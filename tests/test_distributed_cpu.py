"""N > 1 path on CPU: two gloo ranks shard the coupling-term links of every matvec, all-reduce H|psi> once per
Lanczos step and must stay in lock-step (same energies on both ranks, equal to the reference golden)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tests.cpu_backend import CpuBackend, install as cpu_backend_install
        from tests.test_host_logic_cpu import para_from_golden
        from tnalg_b200 import ops
        from tnalg_b200.DMRG_anyH import dmrg_finite_size
        from tnalg_b200 import envs
        envs.EnvCache.shard_min_dim = 4      # exercise the sharded bond moves (in-place all-gather) at chi = 16
        cpu_backend_install(CpuBackend())
        g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', case + '.npz')))
        para = para_from_golden(g)
        np.random.seed(int(g['seed']))
        ob, A, info, para = dmrg_finite_size(para)
        q.put((rank, float(ob['e_per_site'][0]), ob['mz'].reshape(-1).tolist(), [lm.tolist() for lm in A.lm],
               info['flops_executed'] / info['flops_algorithmic']))
    finally:
        dist.destroy_process_group()


def _worker_two_site(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tests.cpu_backend import CpuBackend, install as cpu_backend_install
        from tests.test_two_site_cpu import small_para
        from tnalg_b200 import ops
        from tnalg_b200.DMRG_anyH import dmrg_finite_size_two_site
        cpu_backend_install(CpuBackend())
        para = small_para('xxz', chi=16, sweep_time=6, dt_ob=1, break_tol=1e-13, eigs_tol=1e-14)
        np.random.seed(1)
        ob, A, info, para = dmrg_finite_size_two_site(para, chi_init=2)
        q.put((rank, float(np.ravel(ob['e_per_site'])[0]), [lm.tolist() for lm in A.lm], [int(v) for v in A.virtual_dim]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_two_site_sweep_agree_with_exact_energy():
    from oracle import dmrg_oracle as orc
    from tests.test_two_site_cpu import small_para
    world, port = 2, 31500 + os.getpid() % 2000
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_two_site, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, e0, lm0, vd0), (_, e1, lm1, vd1) = res
    assert e0 == e1 and lm0 == lm1 and vd0 == vd1
    para = small_para('xxz', chi=16)
    exact = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
    assert abs(e0 * para['l'] - exact) < 1e-10 * abs(exact) and max(vd0) == 16


def _scan_paras(tmp):
    """a small (hx) scan of a transverse-field Ising chain, like ScriptRun/DMRG/runDMRGfull.py:27-39"""
    from tnalg_b200 import Parameters as Pm
    out = []
    for hx in (0.2, 0.5, 0.8, 1.1, 1.4):
        para = Pm.generate_parameters_dmrg('chain')
        para.update(l=6, chi=8, jxy=0, jz=1, hx=hx, hz=0, eigs_tol=1e-12, break_tol=1e-13, data_path=tmp)
        out.append(Pm.make_consistent_parameter_dmrg(para))
    return out


def _worker_scan(rank, world, port, tmp, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tests.cpu_backend import CpuBackend, install as cpu_backend_install
        from tnalg_b200 import ops
        from tnalg_b200.DMRG_anyH import run_parameter_scan
        cpu_backend_install(CpuBackend())
        q.put((rank, run_parameter_scan(_scan_paras(tmp), save=True, seed=7)))
    finally:
        dist.destroy_process_group()


def test_parameter_scan_is_distributed_over_ranks_without_term_sharding(tmp_path):
    """independent runs (weak scaling, no data-path collective): every rank reports all runs, energies equal ED, each run
    left its .pr file, and the result equals the serial loop"""
    from oracle import dmrg_oracle as orc
    from tests.cpu_backend import CpuBackend, install as cpu_backend_install
    from tnalg_b200 import BasicFunctionsSJR as bf, ops
    from tnalg_b200.DMRG_anyH import run_parameter_scan
    world, port = 2, 33500 + os.getpid() % 2000
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_scan, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1] and [r[0] for r in res[0]] == [0, 1, 2, 3, 4]
    paras = _scan_paras(str(tmp_path))
    for (n, e, n_sweeps, conv), para in zip(res[0], paras):
        e0 = np.linalg.eigvalsh(orc.dense_hamiltonian(para))[0]
        assert abs(e * para['l'] - e0) < 1e-10 * abs(e0)
        saved = bf.load_pr(os.path.join(str(tmp_path), para['data_exp'] + '.pr'))
        assert abs(float(np.ravel(saved['ob']['e_per_site'])[0]) - e) == 0 and saved['A'].length == para['l']
    old = ops._backend
    cpu_backend_install(CpuBackend())
    try:
        serial = run_parameter_scan(paras, save=False, seed=7)
    finally:
        cpu_backend_install(old)
    assert [r[:2] for r in serial] == [r[:2] for r in res[0]]


@pytest.mark.parametrize('case', ['e2e_j1j2_4x2'])
def test_two_ranks_shard_terms_and_agree(case):
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', case + '.npz')))
    (r0, e0, mz0, lm0, f0), (r1, e1, mz1, lm1, f1) = res
    assert e0 == e1 and mz0 == mz1 and lm0 == lm1            # replicas stay bit-identical
    assert abs(e0 - g['e_per_site'][0]) <= 1e-10 * abs(g['e_per_site'][0])
    assert np.abs(np.array(mz0) - g['mz'].reshape(-1)).max() < 1e-8
    for n, lm in enumerate(lm0):
        assert np.abs(np.array(lm) - g['lm_%d' % n]).max() <= 1e-10 * g['lm_%d' % n].max() + 1e-12
    assert 0.1 < f0 < 0.8 and 0.1 < f1 < 0.8                    # each rank executes only a share of the (merged) link flops


def test_sharded_plans_sum_to_full_operator():
    from tests.cpu_backend import CpuPlan
    rng = np.random.RandomState(0)
    a, d, b = 5, 2, 6
    g = {'HL': rng.randn(a, a), 'HR': rng.randn(b, b), 'M': rng.randn(d, d), 'LS': [rng.randn(a, a) for _ in range(3)],
         'ls_ops': [rng.randn(d, d) for _ in range(3)], 'RS': [rng.randn(b, b) for _ in range(2)],
         'rs_ops': [rng.randn(d, d) for _ in range(2)], 'XL': [rng.randn(a, a) for _ in range(4)],
         'XR': [rng.randn(b, b) for _ in range(4)], 'x_coeff': list(rng.randn(4))}
    x = torch.from_numpy(rng.randn(a, d, b))
    full = CpuPlan((a, d, b), g).matvec(x, 1.0, -0.3).numpy()
    for world in (2, 3, 8):
        parts = sum(CpuPlan((a, d, b), g, r, world).matvec(x, 1.0, -0.3).numpy() for r in range(world))
        assert np.abs(parts - full).max() < 1e-13
