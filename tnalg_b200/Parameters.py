"""Parameter dictionaries of the finite-size DMRG path.  Same entry points, keys and values as the reference's
Parameters.py:6-281,410-446 (boundary: the dict is consumed unchanged by dmrg_finite_size); re-implemented from the
key list in SURVEY.md section 5/8d.  Lattice names are compared with ==, so non-interned strings work too."""
import numpy as np

from . import HamiltonianModule as hm


def common_parameters_dmrg():
    return dict(chi=30, sweep_time=100, if_print_detail=False, tau=1e-4, eigs_tol=1e-5, break_tol=1e-8, is_real=True,
                dt_ob=4, ob_position=0, eigWay=1, isParallel=False, isParallelEnvLMR=False, is_save_op=True,
                data_path='.\\data_dmrg\\')


def _six_ops(spin):
    op = hm.spin_operators(spin)
    return [op['id'], op['sx'], op['sy'], op['sz'], op['su'], op['sd']]


def parameter_dmrg_arbitrary():
    para = dict(spin='half')
    para['op'] = _six_ops('half')
    para['index1'] = [[0, 6], [1, 6], [2, 6]]
    para['coeff1'] = [0.1, 0.1, 0.1]
    para['index2'] = [[0, 1, 3, 3], [1, 2, 3, 3]]
    para['coeff2'] = [1, 1]
    para['lattice'] = 'arbitrary'
    para['data_exp'] = 'Put_here_your_file_name_to_save_data'
    return para


def parameter_dmrg_chain():
    return dict(spin='half', bound_cond='open', l=18, jxy=1, jz=1, hx=0, hz=0, lattice='chain')


def parameter_dmrg_jigsaw():
    return dict(spin='one', bound_cond='open', l=21, jxy=1, jz=1, jxy1=1, jz1=1, hx=0, hz=0, lattice='jigsaw')


def parameter_dmrg_square():
    return dict(bound_cond='open', square_width=4, square_height=4, spin='half', op=_six_ops('half'), jxy=1, jz=1,
                hx=0, hz=0, lattice='square')


def parameter_dmrg_full():
    return dict(spin='half', l=6, jxy=0, jz=1, hx=0.5, hz=0, lattice='full')


def parameter_dmrg_long_range():
    return dict(spin='half', alpha=1, l=6, jxy=0, jz=1, hx=0.5, hz=0, lattice='longRange')


def generate_parameters_dmrg(lattice='chain'):
    makers = {'chain': parameter_dmrg_chain, 'square': parameter_dmrg_square, 'arbitrary': parameter_dmrg_arbitrary,
              'jigsaw': parameter_dmrg_jigsaw, 'full': parameter_dmrg_full, 'longRange': parameter_dmrg_long_range}
    if lattice not in makers:
        raise ValueError('Wrong input of lattice! Set lattice as one of %s' % sorted(makers))
    para = dict(makers[lattice](), **common_parameters_dmrg())
    return make_consistent_parameter_dmrg(para)


def _uniform_field_terms(para):
    """index1/coeff1 and op[6] = -hx*sx - hz*sz on every site (Parameters.py:172-176)"""
    L = para['l']
    para['index1'] = np.stack([np.arange(L), 6 * np.ones(L)], axis=1).astype(int)
    para['coeff1'] = np.ones((L, 1))


def _heisenberg_coeff2(n_bonds, jxy, jz):
    return np.tile(np.array([jxy / 2, jxy / 2, jz], dtype=float), n_bonds).reshape(-1, 1)


def make_consistent_parameter_dmrg(para):
    lattice = para['lattice']
    if lattice in ('chain', 'jigsaw', 'full', 'longRange'):
        para['op'] = _six_ops(para['spin'])
        if lattice == 'jigsaw':
            want_odd = para['bound_cond'] == 'open'
            if (para['l'] % 2 == 0) == want_odd:
                para['l'] += 1
        para['op'].append(-para['hx'] * para['op'][1] - para['hz'] * para['op'][3])
        _uniform_field_terms(para)
    if lattice == 'chain':
        para['positions_h2'] = hm.positions_nearest_neighbor_1d(para['l'], para['bound_cond'])
        para['index2'] = hm.interactions_position2full_index_heisenberg_two_body(para['positions_h2'])
        para['data_exp'] = 'chainN%d_j(%g,%g)_h(%g,%g)_chi%d' % (para['l'], para['jxy'], para['jz'], para['hx'], para['hz'],
                                                                 para['chi']) + para['bound_cond']
        para['coeff2'] = _heisenberg_coeff2(para['positions_h2'].shape[0], para['jxy'], para['jz'])
    elif lattice == 'square':
        para['l'] = para['square_width'] * para['square_height']
        # the reference appends the field operator to whatever para['op'] holds (Parameters.py:190)
        para['op'].append(-para['hx'] * para['op'][1] - para['hz'] * para['op'][3])
        _uniform_field_terms(para)
        para['positions_h2'] = hm.positions_nearest_neighbor_square(para['square_width'], para['square_height'],
                                                                    para['bound_cond'])
        para['index2'] = hm.interactions_position2full_index_heisenberg_two_body(para['positions_h2'])
        para['data_exp'] = 'square(%d,%d)' % (para['square_width'], para['square_height']) + \
            'N%d_j(%g,%g)_h(%g,%g)_chi%d' % (para['l'], para['jxy'], para['jz'], para['hx'], para['hz'], para['chi']) + \
            para['bound_cond']
        para['coeff2'] = _heisenberg_coeff2(para['positions_h2'].shape[0], para['jxy'], para['jz'])
    elif lattice == 'arbitrary':
        para['coeff1'] = np.array(para['coeff1']).reshape(-1, 1)
        para['coeff2'] = np.array(para['coeff2']).reshape(-1, 1)
        para['index1'] = np.array(para['index1'])
        para['index2'] = np.array(para['index2'])
        para['l'] = max(max(para['index1'][:, 0]), max(para['index2'][:, 0]), max(para['index2'][:, 1])) + 1
        para['positions_h2'] = from_index2_to_positions_h2(para['index2'])
        check_continuity_pos_h2(pos_h2=para['positions_h2'])
    elif lattice == 'jigsaw':
        para['positions_h2'] = hm.positions_jigsaw_1d(para['l'], para['bound_cond'])
        para['index2'] = hm.interactions_position2full_index_heisenberg_two_body(para['positions_h2'])
        n_chain = para['l'] - (1 if para['bound_cond'] == 'open' else 0)
        n_bonds = para['positions_h2'].shape[0]
        para['coeff2'] = np.vstack([_heisenberg_coeff2(n_chain, para['jxy'], para['jz']),
                                    _heisenberg_coeff2(n_bonds - n_chain, para['jxy1'], para['jz1'])])
        para['data_exp'] = 'JigsawN%d_j(%g,%g,%g,%g)_h(%g,%g)_chi%d' % (
            para['l'], para['jxy'], para['jz'], para['jxy1'], para['jz1'], para['hx'], para['hz'], para['chi']) + \
            para['bound_cond']
    elif lattice in ('full', 'longRange'):
        para['positions_h2'] = hm.positions_fully_connected(para['l'])
        para['index2'] = hm.interactions_position2full_index_heisenberg_two_body(para['positions_h2'])
        para['coeff2'] = _heisenberg_coeff2(para['positions_h2'].shape[0], para['jxy'], para['jz'])
        if lattice == 'full':
            para['data_exp'] = 'fullConnectedN%d_j(%g,%g)_h(%g,%g)_chi%d' % (para['l'], para['jxy'], para['jz'],
                                                                             para['hx'], para['hz'], para['chi'])
        else:
            dist = np.abs(para['positions_h2'][:, 0] - para['positions_h2'][:, 1]).astype(float) ** para['alpha']
            para['coeff2'] = para['coeff2'] / np.repeat(dist, 3).reshape(-1, 1)
            para['data_exp'] = 'longRangeN%d_j(%g,%g)_h(%g,%g)_chi%d_alpha%g' % (
                para['l'], para['jxy'], para['jz'], para['hx'], para['hz'], para['chi'], para['alpha'])
    else:
        raise ValueError('unknown lattice %r' % (lattice,))
    para['d'] = physical_dim_from_spin(para['spin'])
    para['nh'] = para['index2'].shape[0]
    return para


def from_index2_to_positions_h2(index2):
    """distinct (site1, site2) pairs in lexicographic order (Parameters.py:410-418)"""
    pairs = sorted({(int(r[0]), int(r[1])) for r in np.asarray(index2)})
    return np.array(pairs, dtype=int).reshape(-1, 2)


def check_continuity_pos_h2(pos_h2):
    p0, p1 = int(np.min(pos_h2)), int(np.max(pos_h2))
    if p0 != 0:
        raise SystemExit('The numbering of sites should start with 0, not %d. Please revise the numbering.' % p0)
    missing = [n for n in range(p0 + 1, p1) if n not in pos_h2]
    if missing:
        raise SystemExit('The pos_h2 is expected to contain all numbers from 0 to %d; missing: %s' % (p1, missing))


def physical_dim_from_spin(spin):
    return {'half': 2, 'one': 3}.get(spin, False)


def show_parameters(para):
    from .BasicFunctionsSJR import print_dict
    return print_dict(para, welcome='The parameters are: \n', style_sep=':\n')


# ---- infinite DMRG / TEBD parameter dicts of the library generation (library/Parameters.py:334-369, 430-467) ----
def generate_parameters_infinite_dmrg():
    para = dict(dmrg_type='mpo', spin='half', jxy=0, jz=1, hx=0.5, hz=0, n_site=2, chi=16, sweep_time=1000, tau=1e-4,
                eigs_tol=1e-15, break_tol=2e-10, is_symme_env=False, is_real=True, form='center_ort', dt_ob=10,
                data_path='.\\data_idmrg\\')
    return make_para_consistent_idmrg(para)


def make_para_consistent_idmrg(para):
    if para['dmrg_type'] not in ('mpo', 'white'):
        print("Bad para['d,rg_type']. Set to 'white'")
        para['dmrg_type'] = 'white'
    if para['dmrg_type'] == 'white':
        para['is_symme_env'] = True      # 'white' only suits nearest-neighbour chains with mirror-symmetric environments
    para['model'] = 'heisenberg'
    para['hamilt_index'] = hm.hamiltonian_indexes(para['model'], (para['jxy'], para['jz'], -para['hx'] / 2, -para['hz'] / 2))
    para['d'] = physical_dim_from_spin(para['spin'])
    return para


def generate_parameters_standard_tebd(lattice='chain'):
    para = dict(spin='half', jxy=1, jz=1, hx=0, hz=0, l=12, chi=32, bound_cond='open', tau0=1e-1, dtau=0.1, taut=3, dt_ob=10,
                iterate_time=5000, lattice=lattice, save_mode='final', if_break=True, break_tol=1e-7, data_path='.\\data_tebd\\')
    return make_para_consistent_tebd(para)


def make_para_consistent_tebd(para):
    para['positions_h2'] = hm.positions_nearest_neighbor_1d(para['l'], para['bound_cond'])
    para['num_h2'] = para['positions_h2'].shape[0]
    para['d'] = physical_dim_from_spin(para['spin'])
    para['op'] = _six_ops(para['spin'])
    para['op'].append(-para['hx'] * para['op'][1] - para['hz'] * para['op'][3])
    para['ob_time'] = int(para['iterate_time'] / para['dt_ob'])
    return para
