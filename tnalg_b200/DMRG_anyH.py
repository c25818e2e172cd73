"""dmrg_finite_size: the sweep driver of the reference (DMRG_anyH.py:18-101) on the CUDA path.

Same signature and return value: dmrg_finite_size(para) -> (ob, A, info, para) with
ob keys eb_full, mx, mz, e_per_site, eb, corr_x, corr_z and info keys convergence, t_cost.
Extra (non-reference) info keys: n_sweeps, n_solves, n_matvec, flops_algorithmic, flops_executed.
"""
import time

import numpy as np

from . import Parameters as pm
from .MPSClass import MpsOpenBoundaryClass as Mob

is_debug = False


def sweep_order(length, ob_position):
    """site order of one sweep: right from ob_position+1, back to 0, then up to ob_position-1 (DMRG_anyH.py:47-64)"""
    return list(range(ob_position + 1, length)) + list(range(length - 2, -1, -1)) + list(range(1, ob_position))


def sweep_once(A, para):
    for n in sweep_order(para['l'], para['ob_position']):
        A.update_tensor_eigs(n, para['index1'], para['index2'], para['coeff1'], para['coeff2'], para['tau'],
                             para['is_real'], tol=para['eigs_tol'])


def observe(A, para, ob):
    ob['eb_full'], (ob['mx'], ob['mz']) = A.observe_bond_energy_and_magnetization(para['index2'], para['coeff2'], (1, 3))
    ob['e_per_site'] = (sum(ob['eb_full']) - para['hx'] * sum(ob['mx']) - para['hz'] * sum(ob['mz'])) / A.length
    return ob


def dmrg_finite_size(para=None, quiet=True):
    t_start = time.time()
    info = dict()
    if para is None:
        para = pm.generate_parameters_dmrg()
    say = (lambda *a: None) if quiet else print
    A = Mob(length=para['l'], d=para['d'], chi=para['chi'], way='qr', ini_way='r', operators=para['op'], debug=is_debug,
            is_parallel=para['isParallel'], par_pool=None, is_save_op=para['is_save_op'], eig_way=para['eigWay'],
            is_env_parallel_lmr=para['isParallelEnvLMR'])
    A.shard_terms = bool(para.get('shard_terms', True))   # False: independent runs per rank (parameter scans)
    A.sync_replicas()     # sharded runs work on bit-identical replicas: rank 0's random start goes to every rank
    A.correct_orthogonal_center(para['ob_position'])
    e0_per_site = 0
    info['convergence'] = 1
    info['n_sweeps'] = 0
    ob = dict()
    for t in range(0, para['sweep_time']):
        if_ob = ((t + 1) % para['dt_ob'] == 0) or t == (para['sweep_time'] - 1)
        sweep_once(A, para)
        info['n_sweeps'] = t + 1
        if if_ob:
            observe(A, para, ob)
            info['convergence'] = abs(ob['e_per_site'] - e0_per_site)
            if info['convergence'] < para['break_tol']:
                say('Converged at the %d-th sweep with error = %g of energy per site.' % (t + 1, np.ravel(info['convergence'])[0]))
                break
            say('Convergence error of energy per site = %g' % np.ravel(info['convergence'])[0])
            e0_per_site = ob['e_per_site']
        if t == para['sweep_time'] - 1 and info['convergence'] > para['break_tol']:
            say('Not converged with error = %g of eb per bond' % np.ravel(info['convergence'])[0])
    ob['eb'] = get_bond_energies(ob['eb_full'], para['positions_h2'], para['index2'])
    A.calculate_entanglement_spectrum()
    A.calculate_entanglement_entropy()
    ob['corr_x'] = A.observe_correlators_from_middle(1, 1)
    ob['corr_z'] = A.observe_correlators_from_middle(3, 3)
    info['t_cost'] = time.time() - t_start
    info.update({k: A.stats[k] for k in ('n_solves', 'n_matvec', 'flops_algorithmic', 'flops_executed', 'not_converged')})
    say('Simulation finished in %g seconds' % info['t_cost'])
    A.clean_to_save()
    return ob, A, info, para


def sweep_once_two_site(A, para, chi=None):
    """left-to-right then right-to-left pass of two-site updates with SVD truncation to chi"""
    L = para['l']
    chi = para['chi'] if chi is None else chi
    args = (para['index1'], para['index2'], para['coeff1'], para['coeff2'], para['tau'], para['is_real'])
    for p in range(0, L - 1):
        A.update_two_sites_eigs(p, *args, tol=para['eigs_tol'], chi=chi, to_right=True)
    for p in range(L - 2, -1, -1):
        A.update_two_sites_eigs(p, *args, tol=para['eigs_tol'], chi=chi, to_right=False)


def dmrg_finite_size_two_site(para=None, chi_init=None, quiet=True):
    """Finite-size DMRG with TWO-site updates and SVD truncation to para['chi'] (not in the reference's finite driver; it
    combines the reference's term-summed two-site matvec and SVD truncation, library/MPSClass.py:1676-1707, with the sweep
    of DMRG_anyH.py:18-101).  Same para dict and same (ob, A, info, para) return value as dmrg_finite_size; the MPS starts
    at bond dimension chi_init (default min(chi, 4)) and grows to chi."""
    t_start = time.time()
    if para is None:
        para = pm.generate_parameters_dmrg()
    say = (lambda *a: None) if quiet else print
    chi0 = min(para['chi'], 4) if chi_init is None else chi_init
    A = Mob(length=para['l'], d=para['d'], chi=chi0, way='qr', ini_way='r', operators=para['op'], debug=is_debug,
            is_parallel=para['isParallel'], par_pool=None, is_save_op=para['is_save_op'], eig_way=para['eigWay'],
            is_env_parallel_lmr=para['isParallelEnvLMR'])
    A.shard_terms = bool(para.get('shard_terms', True))
    A.sync_replicas()
    A.correct_orthogonal_center(0)
    info = {'convergence': 1, 'n_sweeps': 0}
    ob, e0 = dict(), 0
    for t in range(0, para['sweep_time']):
        if_ob = ((t + 1) % para['dt_ob'] == 0) or t == (para['sweep_time'] - 1)
        sweep_once_two_site(A, para)
        info['n_sweeps'] = t + 1
        if if_ob:
            observe(A, para, ob)
            info['convergence'] = float(np.ravel(abs(ob['e_per_site'] - e0))[0])
            if info['convergence'] < para['break_tol']:
                say('Converged at the %d-th sweep with error = %g of energy per site.' % (t + 1, info['convergence']))
                break
            e0 = ob['e_per_site']
    ob['eb'] = get_bond_energies(ob['eb_full'], para['positions_h2'], para['index2'])
    A.calculate_entanglement_spectrum()
    A.calculate_entanglement_entropy()
    ob['corr_x'] = A.observe_correlators_from_middle(1, 1)
    ob['corr_z'] = A.observe_correlators_from_middle(3, 3)
    info['t_cost'] = time.time() - t_start
    info.update({k: A.stats[k] for k in ('n_solves', 'n_matvec', 'flops_algorithmic', 'flops_executed', 'not_converged')})
    A.clean_to_save()
    return ob, A, info, para


def dmrg_infinite_size(para=None, A=None, hamilt=None, quiet=True):
    """White-style two-site iDMRG on the CUDA kernels: the driver of algorithms/DMRG_anyH.py:106-176 for dmrg_type='white',
    n_site=2 (the MPO form is not provided).  Returns (A, ob, info) like the reference; ob['eb'] is the energy of the central
    bond at the last observation, ob['eb_history'] every observed value."""
    from .HamiltonianModule import hamiltonian_heisenberg_library
    from .MPSClass import MpsInfinite
    t_start = time.time()
    if para is None:
        para = pm.generate_parameters_infinite_dmrg()
        para.update(dmrg_type='white')
        para = pm.make_para_consistent_idmrg(para)
    say = (lambda *a: None) if quiet else print
    if hamilt is None:
        hamilt = hamiltonian_heisenberg_library(para['spin'], para['jxy'], para['jxy'], para['jz'], para['hx'] / 2, para['hz'] / 2)
    if A is None:
        A = MpsInfinite(para['form'], para['d'], para['chi'], para['d'] ** para['n_site'], n_site=para['n_site'],
                        is_symme_env=para['is_symme_env'], dmrg_type=para['dmrg_type'], hamilt_index=para['hamilt_index'])
    e0, e1, de, history = 0.0, 1.0, 1.0, []
    A.update_ort_tensor_mps('left')
    A.update_bath_onsite()
    A.update_effective_ops()
    for t in range(0, para['sweep_time']):
        A.update_central_tensor((para['tau'], 'full'))
        if t % para['dt_ob'] == 0:
            A.rho_from_central_tensor()
            e1 = A.observe_energy(hamilt)
            history.append(e1)
            say('At the %g-th sweep: Eb = %s' % (t, e1))
            de = abs(e0 - e1) / A.n_site
            if de > para['break_tol']:
                e0 = e1
            else:
                say('Converged with de = %g' % de)
                break
        A.update_ort_tensor_mps('left')
        A.update_bath_onsite()
        A.update_effective_ops()
    ob = {'eb': e1, 'eb_history': np.array(history)}
    info = {'t_cost': time.time() - t_start, 'n_matvec': A.stats['n_matvec'], 'n_solves': A.stats['n_solves']}
    return A, ob, info


def run_parameter_scan(paras, save=True, seed=None, two_site=False):
    """Independent DMRG runs over a list of para dicts, distributed over the ranks of torch.distributed (run n goes to rank
    n mod world; no collective on the data path -- the pattern of the reference's ScriptRun/DMRG/runDMRGfull.py:27-39, 110
    runs over a (theta, alpha) grid, each saved as data_path/data_exp.pr).  Every rank returns the list of
    (index, e_per_site, n_sweeps, convergence) of ALL runs (gathered with all_gather_object); the .pr files are written by
    the rank that owns the run.  Without an initialised process group it is a plain serial loop."""
    import numpy as np
    from . import BasicFunctionsSJR as bf
    try:
        import torch.distributed as dist
        on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:  # pragma: no cover
        dist, on = None, False
    rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    mine = []
    for n, para in enumerate(paras):
        if n % world != rank:
            continue
        para = dict(para)
        para['shard_terms'] = False                      # this rank owns the whole run
        if seed is not None:
            np.random.seed(seed + n)
        ob, A, info, para = (dmrg_finite_size_two_site if two_site else dmrg_finite_size)(para)
        if save:
            bf.save_pr(para.get('data_path', './data_dmrg'), para['data_exp'] + '.pr', (ob, A, info, para), ('ob', 'A', 'info', 'para'))
        mine.append((n, float(np.ravel(ob['e_per_site'])[0]), int(info.get('n_sweeps', 0)), float(np.ravel(info['convergence'])[0])))
    if not on:
        return mine
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    return sorted(r for part in everyone for r in part)


def get_bond_energies(eb_full, positions, index2):
    """sum the per-term energies onto their bond (DMRG_anyH.py:250-258)"""
    positions = np.asarray(positions)
    eb = np.zeros((positions.shape[0], 1))
    for i in range(0, eb_full.size):
        p = (index2[i, 0] == positions[:, 0]) * (index2[i, 1] == positions[:, 1])
        eb[np.nonzero(p)] += eb_full[i]
    return eb


def sort_positions(pos):
    """rows sorted lexicographically (used by Parameters.from_index2_to_positions_h2 in the reference)"""
    pos = np.asarray(pos)
    return pos[np.lexsort((pos[:, 1], pos[:, 0]))]
