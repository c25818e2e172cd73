#!/usr/bin/env python
"""Benchmark of the finite-size DMRG hot path (BASELINE.json metric: sweep wall-time & H.psi matvecs/s at chi=1024,
FP64 tensor-pipe fraction of peak).

    python bench.py --gpus 1 --steps K --warmup W            # our arm, one process per GPU (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the reference path

A step is ONE full DMRG sweep (2L-2 local updates: environments, device Lanczos, gauge move) of the J1-J2 Heisenberg
model on the 6x6 open square lattice at chi = 1024 (BASELINE configs[3], the configuration the metric is quoted on;
it fits one GPU).  `value` is matvecs/s over the timed sweeps with the MPS resident in HBM; `e2e` is the same metric
through the drop-in API with the MPS starting in pinned host memory and the results (observables + tensors) read back
every step.  For N > 1 every local eigenproblem is sharded over the ranks (strong scaling): each rank computes a row slice of
H|psi> for all coupling terms and keeps a slice of the Krylov basis; the exchange step is one NCCL all-gather per Lanczos step
issued by the library (--shard terms: the term-sharded variant with one all-reduce of H|psi>).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (builder, kwargs)
    'j1j2_6x6_chi1024': dict(kind='j1j2', w=6, h=6, chi=1024),
    'heis_chain100_chi256': dict(kind='chain', l=100, chi=256),
    'xxz_chain200_chi512': dict(kind='chain', l=200, chi=512, jxy=1, jz=0.5, hx=0.3, hz=0),
    'j1j2_4x4_chi64': dict(kind='j1j2', w=4, h=4, chi=64),  # quick functional check
    'heis_8x8_chi2048': dict(kind='j1j2', w=8, h=8, chi=2048, j2=0.0),  # BASELINE.json configs[4] (8 GPUs)
}


def build_para(spec, chi=None):
    from tnalg_b200 import HamiltonianModule as hm
    from tnalg_b200 import Parameters as Pm
    spec = dict(spec)
    if chi:
        spec['chi'] = chi
    kind = spec.pop('kind')
    if kind == 'chain':
        para = Pm.generate_parameters_dmrg('chain')
        para.update(spec)
        return Pm.make_consistent_parameter_dmrg(para)
    w, h, j2 = spec.pop('w'), spec.pop('h'), spec.pop('j2', 0.5)
    nn = hm.positions_nearest_neighbor_square(w, h, 'open').astype(int)
    diag = []
    for r in range(h - 1):
        for c in range(w - 1):
            if j2 != 0.0:
                diag.append([r * w + c, (r + 1) * w + c + 1])
                diag.append([r * w + c + 1, (r + 1) * w + c])
    pos = np.vstack([nn, np.array(diag, dtype=int).reshape(-1, 2)])
    jj = np.concatenate([np.ones(nn.shape[0]), j2 * np.ones(len(diag))])
    para = dict(Pm.common_parameters_dmrg())
    op = hm.spin_operators('half')
    L = w * h
    para.update(lattice='arbitrary', spin='half', hx=0, hz=0,
                op=[op['id'], op['sx'], op['sy'], op['sz'], op['su'], op['sd'], np.zeros((2, 2))],
                index1=[[i, 6] for i in range(L)], coeff1=np.ones(L),
                index2=hm.interactions_position2full_index_heisenberg_two_body(pos),
                coeff2=np.stack([jj / 2, jj / 2, jj], axis=1).reshape(-1))
    para.update(spec)
    return Pm.make_consistent_parameter_dmrg(para)


def _json_default(o):
    if isinstance(o, np.generic):
        return o.item()
    raise TypeError(type(o))


def bond_dims(L, d, chi):
    L, d, chi = int(L), int(d), int(chi)   # python ints: d ** L overflows int64 at L = 64
    return [min(d ** n, chi, d ** (L - n)) for n in range(L + 1)]


def site_counts(para):
    """(K_L, K_R, n_x) and the algorithmic flop of one matvec for every site."""
    from tnalg_b200.envs import TermTable
    t = TermTable(para['index1'], para['index2'], para['coeff1'], para['coeff2'], para['op'], para['eigs_tol'])
    dims = bond_dims(para['l'], para['d'], para['chi'])
    out = []
    for p in range(para['l']):
        kl, kr, nx = t.counts(p)
        a, b, d = dims[p], dims[p + 1], int(para['d'])
        out.append(dict(site=p, a=a, b=b, kl=kl, kr=kr, nx=nx, flop=2.0 * a * d * b * (a * (kl + nx) + b * (kr + nx))))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nme)
            except Exception:
                pass
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference path, bounded sample at the widest site, extrapolated by sum F_mv
# ----------------------------------------------------------------------------------------------------------------------
def cpu_matvec_sample(para, counts, budget_s=12.0, max_calls=3):
    """time oracle.apply_handle (the reference's update_tensor_eigs_f_handle_optimized, MPSClass.py:755-776) at the widest
    site on synthetic environment blocks of the real shapes; returns flop rate and the extrapolated sweep matvecs/s."""
    from oracle import dmrg_oracle as orc
    widest = max(counts, key=lambda c: c['flop'])
    a, b, d = widest['a'], widest['b'], para['d']
    rng = np.random.RandomState(0)

    def sym(n):
        g = rng.randn(n, n)
        return (g + g.T) / 2
    env = {}
    ops = [np.real(o) for o in para['op'][:6]]
    if widest['kl'] > 0:
        env['1_0_0'] = sym(a)
        for s in (3, 4, 5)[:widest['kl'] - 1]:
            env['1_%d_0' % s] = sym(a)
    if widest['kr'] > 0:
        env['0_0_1'] = sym(b)
        for s in (3, 4, 5)[:widest['kr'] - 1]:
            env['0_%d_1' % s] = sym(b)
    if widest['nx'] > 0:
        env['1_0_1'] = [[0.5, sym(a), sym(b)] for _ in range(widest['nx'])]
    A = orc.OracleMps(2, d, 2, ops + [np.zeros((d, d))], mps=[np.zeros((1, d, 2)), np.zeros((2, d, 1))])
    x = rng.randn(a * d * b)
    times = []
    t_all = time.time()
    for _ in range(max_calls):
        t0 = time.time()
        A.apply_handle(x, env, (a, d, b), para['tau'])
        times.append(time.time() - t0)
        if time.time() - t_all > budget_s:
            break
    t = min(times)
    flop_rate = widest['flop'] / t
    return dict(t_widest=t, flop_rate=flop_rate, calls=len(times), a=a, b=b, kl=widest['kl'], kr=widest['kr'], nx=widest['nx'])


def cpu_sweep_estimate(counts, sample, matvecs_per_site):
    total_mv = sum(matvecs_per_site)
    t = sum(n * c['flop'] for n, c in zip(matvecs_per_site, counts)) / sample['flop_rate']
    return total_mv / t, t


def default_matvecs_per_site(para, counts):
    """one sweep visits sites 1..L-1 then L-2..0; ~20 handle calls per solve (ncv=20), capped by the
    local dimension"""
    L = para['l']
    order = list(range(1, L)) + list(range(L - 2, -1, -1))
    per = [0] * L
    for p in order:
        n = counts[p]['a'] * para['d'] * counts[p]['b']
        per[p] += min(20, n)
    return per


def config_block(para, workload, world, counts, ncv=20):
    """the `config` object both arms print (identical keys and values for the same workload / N)"""
    widest = max(counts, key=lambda c: c['flop'])
    L = para['l']
    return {'workload': workload, 'L': L, 'chi': para['chi'], 'd': int(para['d']), 'terms': int(para['index2'].shape[0]),
            'eigs_tol': para['eigs_tol'], 'ncv': ncv, 'step': 'one full sweep = %d local updates' % (2 * L - 2),
            'l2': 'inputs larger than L2 (environment blocks of one matvec: %.0f MiB)'
                  % ((widest['kl'] + widest['nx']) * widest['a'] ** 2 * 8 / 2 ** 20 + (widest['kr'] + widest['nx']) * widest['b'] ** 2 * 8 / 2 ** 20),
            'parallelism': 'local problems sharded over %d GPU(s)' % world}


def cpu_local_update_sample(para, counts):
    """ONE local update of the reference path at the widest cut on the host cores, with the oracle port (numpy BLAS + scipy ARPACK):
    the environment work a cached reference does per update -- (L-1) identity transfers of update_all_effective_id
    (MPSClass.py:491-502) plus one transfer per live operator of the bond (get_effective_operators_*, :252-304) -- and
    eigsh(LinearOperator(handle), k=1, which='LM', tol=eigs_tol) on the matvec handle (MPSClass.py:755-776,801-805) with synthetic
    environment blocks of the real shapes.  Returns seconds, handle calls and the flop rate."""
    from scipy.sparse.linalg import LinearOperator, eigsh
    from oracle import dmrg_oracle as orc
    widest = max(counts, key=lambda c: c['flop'])
    a, b, d = widest['a'], widest['b'], para['d']
    rng = np.random.RandomState(0)

    def sym(n):
        g = rng.randn(n, n) / np.sqrt(n)
        return (g + g.T) / 2
    env = {}
    ops = [np.real(o) for o in para['op'][:6]]
    if widest['kl'] > 0:
        env['1_0_0'] = sym(a)
        for s in (3, 4, 5)[:widest['kl'] - 1]:
            env['1_%d_0' % s] = sym(a)
    if widest['kr'] > 0:
        env['0_0_1'] = sym(b)
        for s in (3, 4, 5)[:widest['kr'] - 1]:
            env['0_%d_1' % s] = sym(b)
    if widest['nx'] > 0:
        env['1_0_1'] = [[0.5, sym(a), sym(b)] for _ in range(widest['nx'])]
    A = orc.OracleMps(2, d, 2, ops + [np.zeros((d, d))], mps=[np.zeros((1, d, 2)), np.zeros((2, d, 1))])
    T = rng.randn(a, d, b) / np.sqrt(a * d)
    E = sym(a)
    n_transfers = (para['l'] - 1) + 3 * int(round(np.sqrt(para['l']))) + 1   # identity chain + live operators (3W+1)
    x = rng.randn(a * d * b)
    x /= np.linalg.norm(x)
    calls = [0]

    def handle(v):
        calls[0] += 1
        return A.apply_handle(v, env, (a, d, b), para['tau'])
    t0 = time.time()
    for _ in range(n_transfers):
        orc.transfer_l2r(T, None, E)
    t_env = time.time() - t0
    t1 = time.time()
    eigsh(LinearOperator((x.size, x.size), matvec=handle, dtype=float), k=1, which='LM', v0=x, tol=para['eigs_tol'])
    t_eig = time.time() - t1
    flop = calls[0] * widest['flop'] + n_transfers * 2.0 * a * d * b * (a + b)
    return dict(t=t_env + t_eig, t_env=t_env, t_eig=t_eig, calls=calls[0], n_transfers=n_transfers, flop_rate=flop / (t_env + t_eig),
                mv_rate_flops=calls[0] * widest['flop'] / (t_env + t_eig), a=a, b=b, kl=widest['kl'], kr=widest['kr'], nx=widest['nx'])


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(i.get('num_threads', 1)) for i in threadpool_info()] + [1])
    except Exception:
        return None


def run_reference_arm(args, para, workload):
    """CPU arm: the reference's own path (oracle port: numpy BLAS + scipy ARPACK, all host threads) on the same workload.
    A full chi = 1024 sweep is ~6 minutes of 16-32 cores, so every step times ONE complete local update at the widest cut
    (environment transfers + eigsh on the matvec handle) and the sweep value is extrapolated by sum_p F_mv(p); at most
    `max_measured` steps are really executed so that the run ends within a few minutes."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    counts = site_counts(para)
    per_site = default_matvecs_per_site(para, counts)
    total_mv = sum(per_site)
    sweep_flop = sum(n * c['flop'] for n, c in zip(per_site, counts))
    max_measured = 4
    vals, times, s = [], [], None
    for step in range(min(args.warmup, 1) + min(args.steps, max_measured)):
        s = cpu_local_update_sample(para, counts)
        if step >= min(args.warmup, 1):
            t_sweep = sweep_flop / s['mv_rate_flops']        # matvec flop of the sweep at the rate of a whole local update
            vals.append(total_mv / t_sweep)
            times.append(t_sweep)
    value = float(np.mean(vals))
    threads = blas_threads() or os.cpu_count()
    sample = ('%d of %d steps measured, each ONE full local update of the reference path at the widest cut a=b=%d K_L=%d K_R=%d n_x=%d '
              '(oracle port; %d environment transfers %.2f s + scipy eigsh with %d handle calls %.2f s, %.0f GFLOP/s); sweep matvecs/s '
              'extrapolated by sum_p F_mv(p) over %d matvecs'
              % (len(vals), args.steps, s['a'], s['kl'], s['kr'], s['nx'], s['n_transfers'], s['t_env'], s['calls'], s['t_eig'],
                 s['flop_rate'] / 1e9, total_mv))
    line = {'impl': 'reference', 'metric': 'dmrg_sweep_matvecs_per_s', 'value': value, 'unit': 'matvec/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(times)) * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_block(para, workload, args.gpus, counts),
            'cpu_baseline': {'value': value, 'unit': 'matvec/s', 'cores': threads, 'kind': 'port', 'sample': sample,
                             'extrapolated': True, 'host_cpus': os.cpu_count(),
                             'omp_num_threads_env': os.environ.get('OMP_NUM_THREADS'),
                             'note': 'kind=port: the unmodified reference (oracle/ref_shim.py) needs /root/reference, which does not '
                                     'exist on the GPU box; the port is pinned to it by tests/golden'},
            'e2e': {'value': value, 'unit': 'matvec/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line, default=_json_default), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, dev, n=8192, reps=4):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / best / 1e9  # TFLOP/s


def run_sweeps(torch, dist, A, para, steps, world, dev):
    """time `steps` sweeps with CUDA events between barriers; returns dict of timings and counters (max over ranks)"""
    from tnalg_b200.DMRG_anyH import sweep_once

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    mv0, fa0, fe0 = A.stats['n_matvec'], A.stats['flops_algorithmic'], A.stats['flops_executed']
    A.timing = True
    A.solver_time_ms(), A.phase_times_ms()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.nvtx.range_push('timed')   # ncu --nvtx --nvtx-include "timed/" profiles exactly the timed sweeps
    for _ in range(steps):
        sweep_once(A, para)
    torch.cuda.nvtx.range_pop()
    e1.record()
    barrier()
    t_ms = max_over_ranks(e0.elapsed_time(e1))
    solver_ms = max_over_ranks(A.solver_time_ms())
    ph = A.phase_times_ms()
    ph = {k: max_over_ranks(v) for k, v in ph.items()}
    A.timing = False
    return dict(t_ms=t_ms, solver_ms=solver_ms, gauge_ms=ph['gauge'], env_ms=ph['env'], n_mv=A.stats['n_matvec'] - mv0,
                f_alg=A.stats['flops_algorithmic'] - fa0, f_exe=A.stats['flops_executed'] - fe0, barrier=barrier,
                max_over_ranks=max_over_ranks)


def _debug_sweep(torch, A, para):
    """one sweep with finite-value checks after every primitive (debugging aid, TN_BENCH_DEBUG=1)"""
    from tnalg_b200 import ops
    from tnalg_b200.DMRG_anyH import sweep_order
    be = ops.backend()
    fin = lambda t: bool(torch.isfinite(t).all())  # noqa: E731
    qr_real, env_real, lz_real, mp_real = be.qr_tensor, be.env_update, be.lanczos, be.mode_product

    def qr_chk(T, l2r):
        Q, R = qr_real(T, l2r)
        torch.cuda.synchronize()
        if not (fin(Q) and fin(R)):
            Q2, R2 = qr_real(T, l2r)
            raise RuntimeError('QR non-finite: T %s finite=%s l2r=%s |T|max=%g; repeat finite=%s ptr=%x'
                               % (tuple(T.shape), fin(T), l2r, float(T.abs().max()), fin(Q2) and fin(R2), T.data_ptr()))
        return Q, R

    def mp_chk(T, mat, bond):
        out = mp_real(T, mat, bond)
        if not fin(out):
            raise RuntimeError('mode_product non-finite: T %s %s mat %s %s bond %d' % (tuple(T.shape), fin(T), tuple(mat.shape), fin(mat), bond))
        return out

    def env_chk(direction, T, outputs):
        res = env_real(direction, T, outputs)
        for r in res:
            if not fin(r):
                raise RuntimeError('env_update non-finite: T %s finite=%s inputs finite=%s' % (tuple(T.shape), fin(T), [[E is None or fin(E) for E, _ in ln] for ln in outputs]))
        return res

    def lz_chk(plan, tau, v0, tol, **kw):
        if not fin(v0):
            raise RuntimeError('Lanczos start vector non-finite %s' % (plan.shape,))
        y = plan.matvec(v0.reshape(plan.shape))
        if not fin(y):
            raise RuntimeError('matvec non-finite %s' % (plan.shape,))
        return lz_real(plan, tau, v0, tol, **kw)
    be.qr_tensor, be.env_update, be.lanczos, be.mode_product = qr_chk, env_chk, lz_chk, mp_chk
    try:
        for n in sweep_order(para['l'], para['ob_position']):
            try:
                A.update_tensor_eigs(n, para['index1'], para['index2'], para['coeff1'], para['coeff2'], para['tau'], para['is_real'], tol=para['eigs_tol'])
            except Exception as e:
                print('DEBUG: failed at site %d: %s' % (n, e), file=sys.stderr, flush=True)
                raise
    finally:
        be.qr_tensor, be.env_update, be.lanczos, be.mode_product = qr_real, env_real, lz_real, mp_real


def quick_workload(torch, dist, name, world, dev, warmup=1, steps=1):
    """a short run of another BASELINE configuration (sweep wall-time, matvecs/s) for the `other_workloads` block"""
    from tnalg_b200.DMRG_anyH import sweep_once
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    para = build_para(WORKLOADS[name])
    np.random.seed(0)
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.sync_replicas()
    A.correct_orthogonal_center(para['ob_position'])
    if os.environ.get('TN_BENCH_DEBUG'):
        _debug_sweep(torch, A, para)
    for _ in range(warmup):
        sweep_once(A, para)
    r = run_sweeps(torch, dist, A, para, steps, world, dev)
    out = {'L': para['l'], 'chi': para['chi'], 'ms_per_sweep': r['t_ms'] / steps, 'matvecs_per_s': r['n_mv'] / (r['t_ms'] * 1e-3),
           'solver_ms_per_sweep': r['solver_ms'] / steps, 'gauge_qr_ms_per_sweep': r['gauge_ms'] / steps,
           'env_update_ms_per_sweep': r['env_ms'] / steps, 'algorithmic_tflops_sweep': r['f_alg'] / (r['t_ms'] * 1e-3) / 1e12,
           'warmup': warmup, 'steps': steps, 'n_gpus': world}
    del A
    torch.cuda.empty_cache()
    return out


def run_ours(args, para, workload):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from tnalg_b200 import ops
    from tnalg_b200.DMRG_anyH import observe, sweep_once
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    be = ops.backend()
    if args.shard:
        be.shard_mode = args.shard
    counts = site_counts(para)
    L, d, chi = para['l'], para['d'], para['chi']

    np.random.seed(0)
    A = MpsOpenBoundaryClass(L, d, chi, operators=para['op'], is_save_op=True, eig_way=1)
    A.sync_replicas()
    A.correct_orthogonal_center(para['ob_position'])
    for _ in range(args.warmup):
        sweep_once(A, para)
    # ---- timed region: K sweeps, MPS resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    be.lib.tn_launch_count_reset()
    comm = be.comm()
    coll0 = comm.collectives() if comm is not None else 0
    peer0 = comm.peer_collectives() if comm is not None else 0
    r = run_sweeps(torch, dist, A, para, args.steps, world, dev)
    barrier, max_over_ranks = r['barrier'], r['max_over_ranks']
    t_ms, solver_ms, n_mv, f_alg, f_exe = r['t_ms'], r['solver_ms'], r['n_mv'], r['f_alg'], r['f_exe']
    launches = be.launch_count()
    collectives = (comm.collectives() - coll0) if comm is not None else 0
    peer_collectives = (comm.peer_collectives() - peer0) if comm is not None else 0
    clocks = sampler.stop() if rank == 0 else None
    value = n_mv / (t_ms * 1e-3)
    sharding = A.last_eig.get('sharding', 'none')
    not_conv = A.stats['not_converged']

    skip = set(os.environ.get('TN_BENCH_SKIP', '').split(','))   # debugging aid: leave out legs of the run
    # ---- e2e: host buffers in, results out, through the drop-in API ----
    host = [t.cpu().pin_memory() for t in A.mps] if 'pin' not in skip else []
    center = A.center
    h2d = sum(t.numel() * 8 for t in host)
    e2e_mv0 = A.stats['n_matvec']
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    e2e_parts = {'load (H2D)': 0.0, 'sweep (incl. environment rebuild)': 0.0, 'observe (D2H of the observables)': 0.0, 'save (D2H of the tensors)': 0.0}

    def lap(key, t_prev):
        torch.cuda.synchronize()
        now = time.perf_counter()
        e2e_parts[key] += (now - t_prev) * 1e3 / max(args.steps, 1)
        return now

    def e2e_step_untimed():
        A.load_tensors(host, center)
        sweep_once(A, para)
        observe(A, para, {})
        A.clean_to_save()

    if not args.no_e2e:
        e2e_step_untimed()                                         # warm-up of the end-to-end path (page-locked staging buffers)
        e2e_mv0 = A.stats['n_matvec']
        barrier()
        t0 = time.perf_counter()
    for _ in range(0 if args.no_e2e else args.steps):
        B = A                                                      # same object, state re-loaded from the host copy
        tp = time.perf_counter()
        B.load_tensors(host, center)                               # H2D from pinned memory; drops every cached block
        tp = lap('load (H2D)', tp)
        sweep_once(B, para)
        tp = lap('sweep (incl. environment rebuild)', tp)
        ob = observe(B, para, {})                                  # D2H of the energy / magnetisation read-back
        tp = lap('observe (D2H of the observables)', tp)
        B.clean_to_save()                                          # D2H of the tensors (what dmrg_finite_size returns)
        tp = lap('save (D2H of the tensors)', tp)
        d2h = sum(t.nbytes for t in B.mps) + sum(np.asarray(v).nbytes for v in ob.values())
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = (A.stats['n_matvec'] - e2e_mv0) / e2e_s if not args.no_e2e else None

    # ---- roofline of the dominant kernel (chain GEMM inside the matvec) at the widest site ----
    widest = max(counts, key=lambda c: c['flop'])
    p = widest['site']
    A.correct_orthogonal_center(p)
    a_w, b_w = A.mps[p].shape[0], A.mps[p].shape[2]
    rows = be.shard_rows(a_w, d, b_w, rank, world) if world > 1 else None
    plan = A.effective_hamiltonian_plan(p, para['index1'], para['index2'], para['coeff1'], para['coeff2'], tol=para['eigs_tol'],
                                        rank=rank if (world > 1 and rows is None) else 0, world=world if rows is None else 1, rows=rows)
    x = A.mps[p].clone()
    y = torch.empty_like(x) if rows is None else be.empty(rows[1], d, b_w)
    for _ in range(3 if 'roof' not in skip else 0):
        plan.matvec(x, 0.0, 1.0, out=y)
    torch.cuda.synchronize()
    reps = 8 if 'roof' not in skip else 1
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(reps):
        plan.matvec(x, 0.0, 1.0, out=y)
    r1.record()
    torch.cuda.synchronize()
    mv_ms = r0.elapsed_time(r1) / reps
    achieved = plan.flops_algorithmic / world / (mv_ms * 1e-3) / 1e12
    executed_tf = plan.flops_executed / (mv_ms * 1e-3) / 1e12
    executed_flop = plan.flops_executed
    uses_tma = plan.uses_tma
    plan.destroy()
    # one QR gauge move at the widest matricisation (own Householder kernels)
    Tq = A.mps[p]
    for _ in range(2 if 'qr' not in skip else 0):
        be.qr_tensor(Tq, True)
    torch.cuda.synchronize()
    r0.record()
    for _ in range(4 if 'qr' not in skip else 0):
        be.qr_tensor(Tq, True)
    r1.record()
    torch.cuda.synchronize()
    qr_ms = r0.elapsed_time(r1) / 4
    peak = measure_fp64_peak(torch, dev) if 'peak' not in skip else 35.0
    dmma_peak = be.dmma_peak_tflops() if 'dmma' not in skip else 37.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get(workload), tj.get('source')
        except Exception:
            traffic = None

    other = {}
    if not args.no_other and args.workload == 'j1j2_6x6_chi1024':
        del x, y
        torch.cuda.empty_cache()
        # the headline above is complete at this point: a failure in an extra workload is reported in its entry, not raised
        # (every rank runs the same deterministic code, so a failure is the same on all of them)
        if world == 1:
            for name in ('heis_chain100_chi256', 'xxz_chain200_chi512'):
                try:
                    other[name] = quick_workload(torch, dist, name, world, dev, warmup=1, steps=1)
                except Exception as e:
                    other[name] = {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}
        elif world == 8 and not args.no_cfg5:
            A.clean_to_save()
            del A
            torch.cuda.empty_cache()
            try:
                other['heis_8x8_chi2048'] = quick_workload(torch, dist, 'heis_8x8_chi2048', world, dev, warmup=1, steps=1)
            except Exception as e:
                other['heis_8x8_chi2048'] = {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- CPU baseline (rank 0, N = 1 only): one full local update of the oracle port at the widest cut ----
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=os.cpu_count())
        except Exception:
            pass
        s = cpu_local_update_sample(para, counts)
        per_site = default_matvecs_per_site(para, counts)
        sweep_flop = sum(n * c['flop'] for n, c in zip(per_site, counts))
        t = sweep_flop / s['mv_rate_flops']
        cpu = {'value': sum(per_site) / t, 'unit': 'matvec/s', 'cores': blas_threads() or os.cpu_count(), 'kind': 'port',
               'sample': ('ONE full local update of the reference path at the widest cut a=b=%d K_L=%d K_R=%d n_x=%d (oracle port: %d '
                          'environment transfers %.2f s + scipy eigsh with %d handle calls %.2f s, %.0f GFLOP/s); sweep matvecs/s '
                          'extrapolated by sum_p F_mv(p)' % (s['a'], s['kl'], s['kr'], s['nx'], s['n_transfers'], s['t_env'], s['calls'],
                                                             s['t_eig'], s['flop_rate'] / 1e9)),
               'sweep_s_extrapolated': t, 'host_cpus': os.cpu_count()}
    other_ms = t_ms - solver_ms - r['gauge_ms'] - r['env_ms']
    line = {
        'metric': 'dmrg_sweep_matvecs_per_s', 'value': value, 'unit': 'matvec/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': t_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_block(para, workload, world, counts),
        'sweep': {'matvecs_per_sweep': n_mv / args.steps, 'solver_ms_per_sweep': solver_ms / args.steps,
                  'algorithmic_tflops_sweep': f_alg / (t_ms * 1e-3) / 1e12, 'algorithmic_tflops_solver': f_alg / (solver_ms * 1e-3) / 1e12,
                  'executed_tflops_solver_per_gpu': f_exe / (solver_ms * 1e-3) / 1e12, 'not_converged': not_conv,
                  'sharding': sharding, 'collectives_per_sweep': collectives / args.steps,
                  'peer_window_collectives_per_sweep': peer_collectives / args.steps,
                  'phases_ms_per_sweep': {'solver (Lanczos: matvec + collectives + vector ops)': solver_ms / args.steps,
                                          'gauge moves (Householder QR + absorb)': r['gauge_ms'] / args.steps,
                                          'environment updates': r['env_ms'] / args.steps,
                                          'plan build / merges / host gap': other_ms / args.steps}},
        'roofline': {'bound': 'tensor', 'kernel': 'chain_gemm_tma_kernel (left + right stage launches of one matvec at the widest site)',
                     'achieved': executed_tf, 'peak': peak, 'unit': 'TFLOP/s', 'frac': executed_tf / peak, 'traffic': traffic,
                     'traffic_source': traffic_src,
                     'flop_per_matvec_executed': executed_flop, 'flop_per_matvec_reference_grouping': widest['flop'],
                     'achieved_reference_grouping': achieved, 'tma_stage_mask': uses_tma,
                     'note': 'achieved = flop the two launches execute (this rank) / CUDA-event time. Crossing terms that share an operator are summed '
                             'before the GEMM (same H_eff; e.g. n_x 42 -> 15 links at the widest 6x6 J1-J2 site), so the kernels execute fewer flop than the '
                             'reference grouping of SURVEY 8d (K_L, K_R, n_x); counted in reference-grouping flop the same matvec runs at '
                             'achieved_reference_grouping TFLOP/s per GPU, which may exceed the hardware peak',
                     'peak_source': 'cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)',
                     'dmma_issue_peak_measured': dmma_peak, 'frac_of_dmma_issue_peak': executed_tf / dmma_peak,
                     'ms_per_matvec': mv_ms, 'site': p, 'a': widest['a'], 'b': widest['b'], 'K_L': widest['kl'], 'K_R': widest['kr'],
                     'n_x': widest['nx'], 'rows_of_this_rank': None if rows is None else list(rows),
                     'qr_ms_2chi_x_chi': qr_ms},
        'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': 'matvec/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_s * 1e3 / args.steps, 'phases_ms_per_step': e2e_parts},
        'gpu_launches': launches, 'clocks': clocks, 'other_workloads': other,
    }
    print(json.dumps(line, default=_json_default), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='j1j2_6x6_chi1024', choices=sorted(WORKLOADS))
    ap.add_argument('--chi', type=int, default=0, help='override the bond dimension (functional checks)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-e2e', action='store_true', help='skip the e2e leg (profiling runs)')
    ap.add_argument('--no-other', action='store_true', help='skip the other_workloads block (chi=256 / chi=512 chains at N=1)')
    ap.add_argument('--no-cfg5', action='store_true', help='N=8: skip the 8x8 chi=2048 workload (BASELINE configs[4]) in other_workloads')
    ap.add_argument('--shard', default='', choices=['', 'rows', 'terms'], help="multi-GPU decomposition (default: 'rows')")
    args = ap.parse_args()
    if args.impl == 'reference':
        if int(os.environ.get('RANK', '0')) != 0:
            return                                   # rank 0 alone runs the CPU arm
        # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core: re-exec once with the BLAS pools opened up
        n_cpu = str(os.cpu_count() or 1)
        if os.environ.get('TNALG_BENCH_REEXEC') != '1' and any(os.environ.get(k) != n_cpu for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS')):
            env = dict(os.environ, OMP_NUM_THREADS=n_cpu, OPENBLAS_NUM_THREADS=n_cpu, MKL_NUM_THREADS=n_cpu, TNALG_BENCH_REEXEC='1')
            sys.stdout.flush()
            os.execve(sys.executable, [sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env)
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=int(n_cpu))
        except Exception:
            pass
    para = build_para(WORKLOADS[args.workload], args.chi)
    workload = args.workload if not args.chi else '%s@chi%d' % (args.workload, args.chi)
    if args.impl == 'reference':
        run_reference_arm(args, para, workload)
    else:
        run_ours(args, para, workload)


if __name__ == '__main__':
    main()
