"""tebd_standard: imaginary-time TEBD with one two-site Hamiltonian on every bond of para['positions_h2'], the driver of the
reference's algorithms/TEBDalgo.py:11-90 on the CUDA path (gate application through tn_apply_site_op, truncation through the
Jacobi SVD, observables through the batched environment chains).  Same para dict (Parameters.generate_parameters_standard_tebd),
same return value (mps, ob, para)."""
import time

import numpy as np
import scipy.linalg as la

from . import Parameters as pm
from .HamiltonianModule import hamiltonian_heisenberg_library
from .MPSClass import MpsStandardTEBD


def split_gate(hamilt, tau, d):
    """exp(-tau h) as two three-index tensors (TEBDalgo.py:38-46): gates[0] (d, dd, d) acts on the left site and opens the bond
    index dd = d*d, gates[1] (d, dd, d) closes it on the right site"""
    u2 = la.expm(-tau * np.real(hamilt))
    g0, lm0, g1 = np.linalg.svd(u2.reshape(d, d, d, d).transpose(0, 2, 1, 3).reshape(d * d, d * d))
    g0 = (g0.dot(np.diag(np.sqrt(lm0)))).reshape(d, d, d * d).transpose(0, 2, 1)
    g1 = (np.diag(np.sqrt(lm0)).dot(g1)).reshape(d * d, d, d).transpose(1, 0, 2)
    return [g0, g1]


def tebd_standard(para=None, quiet=True):
    start = time.time()
    if para is None:
        para = pm.generate_parameters_standard_tebd()
    say = (lambda *a: None) if quiet else print
    ob = {k: [] for k in ('lm', 'ent', 'mx', 'mz', 'eb', 'e_site')}
    mps = MpsStandardTEBD(para['l'], para['d'], para['chi'], para['spin'], evolve_way='gates')
    mps.correct_orthogonal_center(0)
    mps.norm_mps(True)
    hamilt = hamiltonian_heisenberg_library(para['spin'], para['jxy'], para['jxy'], para['jz'], para['hx'] / 2, para['hz'] / 2)
    n_now, time_now, e0 = 0, 0.0, 100.0

    def observe():
        mps.calculate_entanglement_spectrum()
        mps.calculate_entanglement_entropy()
        ob['lm'].append([np.array(x) for x in mps.lm])
        ob['ent'].append(np.array(mps.ent))
        ob['mx'].append(mps.observe_magnetization(1))
        ob['mz'].append(mps.observe_magnetization(3))
        ob['eb'].append(mps.observe_bond_energy_from_jxyz(para['positions_h2'], para['jxy'], para['jxy'], para['jz']))
        ob['e_site'].append(float((np.sum(ob['eb'][-1]) + np.sum(ob['mx'][-1]) * para['hx'] + np.sum(ob['mz'][-1]) * para['hz']) / para['l']))

    for nt in range(0, para['taut']):
        tau = para['tau0'] * (para['dtau'] ** nt)
        say('Now tau = ' + str(tau))
        gates = split_gate(hamilt, tau, para['d'])
        for t in range(para['iterate_time']):
            for nh in range(para['num_h2']):
                p1, p2 = int(para['positions_h2'][nh, 0]), int(para['positions_h2'][nh, 1])
                mps.evolve_gate_tebd(p1, p2, gates)
                mps.truncate_mps_tebd(p1, p2)
            mps.norm_mps(True)
            time_now += tau
            if para['save_mode'] == 'all' and (t % para['dt_ob']) == 0:
                observe()
                n_now += 1
            elif para['if_break'] is True:
                e1 = float(np.sum(mps.observe_bond_energy_from_jxyz(para['positions_h2'], para['jxy'], para['jxy'], para['jz'])) / para['l'])
                if (abs(e0 - e1) < para['break_tol'] and (t % para['dt_ob']) == 0) or t == para['iterate_time'] - 1:
                    observe()
                    say('Converged with E = %s; Convergence = %g' % (ob['e_site'][-1], abs(e0 - e1)))
                    n_now += 1
                    break
                elif (t % para['dt_ob']) == 0:
                    e0 = e1
    info_t = time.time() - start
    say('TEBD finished with %g s' % info_t)
    mps.clean_to_save()
    return mps, ob, para
