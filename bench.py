#!/usr/bin/env python
"""Benchmark of the finite-size DMRG hot path (BASELINE.json metric: sweep wall-time & H.psi matvecs/s at chi=1024,
FP64 tensor-pipe fraction of peak).

    python bench.py --gpus 1 --steps K --warmup W            # our arm, one process per GPU (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the reference path

A step is ONE full DMRG sweep (2L-2 local updates: environments, device Lanczos, gauge move) of the J1-J2 Heisenberg
model on the 6x6 open square lattice at chi = 1024 (BASELINE configs[3], the configuration the metric is quoted on;
it fits one GPU).  `value` is matvecs/s over the timed sweeps with the MPS resident in HBM; `e2e` is the same metric
through the drop-in API with the MPS starting in pinned host memory and the results (observables + tensors) read back
every step.  For N > 1 the coupling-term links of every matvec are sharded over the ranks (strong scaling) with one
NCCL all-reduce of H|psi> per Lanczos step.
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


def run_reference_arm(args, para, workload):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    counts = site_counts(para)
    threads = os.cpu_count()
    per_site = default_matvecs_per_site(para, counts)
    vals, times = [], []
    for step in range(args.warmup + args.steps):
        s = cpu_matvec_sample(para, counts, budget_s=8.0, max_calls=1)
        v, t = cpu_sweep_estimate(counts, s, per_site)
        if step >= args.warmup:
            vals.append(v)
            times.append(t)
    value = float(np.mean(vals))
    sample = ('per step: 1 call of the reference matvec handle (oracle port, numpy BLAS) at the widest site '
              'a=b=%d K_L=%d K_R=%d n_x=%d; sweep matvecs/s extrapolated by sum_p F_mv(p) over %d matvecs'
              % (s['a'], s['kl'], s['kr'], s['nx'], sum(per_site)))
    line = {'impl': 'reference', 'metric': 'dmrg_sweep_matvecs_per_s', 'value': value, 'unit': 'matvec/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(times)) * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload, 'L': para['l'], 'chi': para['chi'], 'd': int(para['d']), 'terms': int(para['index2'].shape[0]),
                       'eigs_tol': para['eigs_tol'], 'ncv': 20, 'step': 'one full sweep = %d local updates' % (2 * para['l'] - 2),
                       'extrapolated': True},
            'cpu_baseline': {'value': value, 'unit': 'matvec/s', 'cores': threads, 'kind': 'port', 'sample': sample,
                             'gflops_widest': s['flop_rate'] / 1e9},
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
    counts = site_counts(para)
    L, d, chi = para['l'], para['d'], para['chi']

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

    np.random.seed(0)
    A = MpsOpenBoundaryClass(L, d, chi, operators=para['op'], is_save_op=True, eig_way=1)
    A.correct_orthogonal_center(para['ob_position'])
    for _ in range(args.warmup):
        sweep_once(A, para)
    # ---- timed region: K sweeps, MPS resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    mv0, fa0, fe0 = A.stats['n_matvec'], A.stats['flops_algorithmic'], A.stats['flops_executed']
    be.lib.tn_launch_count_reset()
    A.timing = True
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.nvtx.range_push('timed')   # ncu --nvtx --nvtx-include "timed/" profiles exactly the timed sweeps
    for _ in range(args.steps):
        sweep_once(A, para)
    torch.cuda.nvtx.range_pop()
    e1.record()
    barrier()
    t_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = be.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    solver_ms = max_over_ranks(A.solver_time_ms())
    A.timing = False
    n_mv = A.stats['n_matvec'] - mv0
    f_alg = A.stats['flops_algorithmic'] - fa0
    f_exe = A.stats['flops_executed'] - fe0
    value = n_mv / (t_ms * 1e-3)

    # ---- e2e: host buffers in, results out, through the drop-in API ----
    host = [t.cpu().pin_memory() for t in A.mps]
    center = A.center
    h2d = sum(t.numel() * 8 for t in host)
    e2e_mv0 = A.stats['n_matvec']
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(0 if args.no_e2e else args.steps):
        B = A                                                      # same object, state re-loaded from the host copy
        B.load_tensors(host, center)                               # H2D from pinned memory; drops every cached block
        sweep_once(B, para)
        ob = observe(B, para, {})                                  # D2H of the energy / magnetisation read-back
        B.clean_to_save()                                          # D2H of the tensors (what dmrg_finite_size returns)
        d2h = sum(t.nbytes for t in B.mps) + sum(np.asarray(v).nbytes for v in ob.values())
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = (A.stats['n_matvec'] - e2e_mv0) / e2e_s if not args.no_e2e else None

    # ---- roofline of the dominant kernel (chain GEMM inside the matvec) at the widest site ----
    widest = max(counts, key=lambda c: c['flop'])
    p = widest['site']
    A.correct_orthogonal_center(p)
    plan = A.effective_hamiltonian_plan(p, para['index1'], para['index2'], para['coeff1'], para['coeff2'], tol=para['eigs_tol'],
                                        rank=rank if world > 1 else 0, world=world)
    x = A.mps[p].clone()
    y = torch.empty_like(x)
    for _ in range(3):
        plan.matvec(x, 0.0, 1.0, out=y)
    torch.cuda.synchronize()
    reps = 8
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(reps):
        plan.matvec(x, 0.0, 1.0, out=y)
    r1.record()
    torch.cuda.synchronize()
    mv_ms = r0.elapsed_time(r1) / reps
    achieved = plan.flops_algorithmic / world / (mv_ms * 1e-3) / 1e12 if world > 1 else plan.flops_algorithmic / (mv_ms * 1e-3) / 1e12
    executed_tf = plan.flops_executed / (mv_ms * 1e-3) / 1e12
    executed_flop = plan.flops_executed
    uses_tma = plan.uses_tma
    plan.destroy()
    peak = measure_fp64_peak(torch, dev)
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get(workload)
        except Exception:
            traffic = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        s = cpu_matvec_sample(para, counts, budget_s=15.0, max_calls=3)
        per_site = default_matvecs_per_site(para, counts)
        v, t = cpu_sweep_estimate(counts, s, per_site)
        cpu = {'value': v, 'unit': 'matvec/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': ('%d calls of the reference matvec handle (oracle port, numpy BLAS) at the widest site a=b=%d K_L=%d '
                          'K_R=%d n_x=%d (%.2f s each, %.1f GFLOP/s); sweep matvecs/s extrapolated by sum_p F_mv(p)'
                          % (s['calls'], s['a'], s['kl'], s['kr'], s['nx'], s['t_widest'], s['flop_rate'] / 1e9)),
               'sweep_s_extrapolated': t}
    line = {
        'metric': 'dmrg_sweep_matvecs_per_s', 'value': value, 'unit': 'matvec/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': t_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload, 'L': L, 'chi': chi, 'd': d, 'terms': int(para['index2'].shape[0]),
                   'eigs_tol': para['eigs_tol'], 'ncv': A.lanczos_ncv, 'step': 'one full sweep = %d local updates' % (2 * L - 2),
                   'l2': 'inputs larger than L2 (environment blocks of one matvec: %.0f MiB)'
                         % ((widest['kl'] + widest['nx']) * widest['a'] ** 2 * 8 / 2 ** 20 + (widest['kr'] + widest['nx']) * widest['b'] ** 2 * 8 / 2 ** 20),
                   'parallelism': 'terms sharded over %d GPU(s)' % world},
        'sweep': {'matvecs_per_sweep': n_mv / args.steps, 'solver_ms_per_sweep': solver_ms / args.steps,
                  'algorithmic_tflops_sweep': f_alg / (t_ms * 1e-3) / 1e12, 'algorithmic_tflops_solver': f_alg / (solver_ms * 1e-3) / 1e12,
                  'executed_tflops_solver': f_exe / (solver_ms * 1e-3) / 1e12, 'not_converged': A.stats['not_converged']},
        'roofline': {'bound': 'tensor', 'kernel': 'chain_gemm_tma_kernel (left + right stage launches of one matvec at the widest site)',
                     'achieved': executed_tf, 'peak': peak, 'unit': 'TFLOP/s', 'frac': executed_tf / peak, 'traffic': traffic,
                     'flop_per_matvec_executed': executed_flop, 'flop_per_matvec_reference_grouping': widest['flop'],
                     'achieved_reference_grouping': achieved, 'tma_stage_mask': uses_tma,
                     'note': 'achieved = flop the two launches execute / CUDA-event time. Crossing terms that share an operator are summed '
                             'before the GEMM (same H_eff; e.g. n_x 42 -> 15 links at the widest 6x6 J1-J2 site), so the kernels execute fewer flop than the '
                             'reference grouping of SURVEY 8d (K_L, K_R, n_x); counted in reference-grouping flop the same matvec runs at '
                             'achieved_reference_grouping TFLOP/s, which may exceed the hardware peak',
                     'peak_source': 'cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry); '
                                    'DMMA issue peak 37.09 TFLOP/s (profiles/r01_fp64_peaks.txt)',
                     'ms_per_matvec': mv_ms, 'site': p, 'a': widest['a'], 'b': widest['b'], 'K_L': widest['kl'], 'K_R': widest['kr'],
                     'n_x': widest['nx']},
        'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': 'matvec/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_s * 1e3 / args.steps},
        'gpu_launches': launches, 'clocks': clocks,
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
    args = ap.parse_args()
    para = build_para(WORKLOADS[args.workload], args.chi)
    workload = args.workload if not args.chi else '%s@chi%d' % (args.workload, args.chi)
    if args.impl == 'reference':
        run_reference_arm(args, para, workload)
    else:
        run_ours(args, para, workload)


if __name__ == '__main__':
    main()
