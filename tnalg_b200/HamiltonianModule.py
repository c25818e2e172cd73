"""Model tables consumed by the DMRG path: spin operators, lattice bond lists and the Heisenberg term expansion.
Same function names and return values as the reference's HamiltonianModule.py:8-44,151-261 (boundary, unchanged
semantics); written from the specification in SURVEY.md, not from the reference source."""
import numpy as np


def spin_operators(spin):
    """dict id/sx/sy/sz/su/sd.  'half': S=1/2 (sx, sz carry the 1/2, su/sd are the bare ladder matrices);
    'one': S=1 with su = sx + i sy, sd = sx - i sy (HamiltonianModule.py:8-44)."""
    op = dict()
    if spin == 'half':
        op['id'] = np.eye(2)
        op['sx'] = np.array([[0., 0.5], [0.5, 0.]])
        op['sy'] = np.array([[0., 0.5j], [-0.5j, 0.]])
        op['sz'] = np.array([[0.5, 0.], [0., -0.5]])
        op['su'] = np.array([[0., 1.], [0., 0.]])
        op['sd'] = np.array([[0., 0.], [1., 0.]])
    elif spin == 'one':
        r = 2 ** 0.5
        op['id'] = np.eye(3)
        op['sx'] = np.array([[0., 1., 0.], [1., 0., 1.], [0., 1., 0.]]) / r
        op['sy'] = np.array([[0., -1j, 0.], [1j, 0., -1j], [0., 1j, 0.]]) / r
        op['sz'] = np.diag([1., 0., -1.])
        op['su'] = np.real(op['sx'] + 1j * op['sy'])
        op['sd'] = np.real(op['sx'] - 1j * op['sy'])
    else:
        raise ValueError("spin must be 'half' or 'one'")
    return op


def positions_nearest_neighbor_1d(l, bound_cond='open'):
    """[[n, n+1]] for the open chain, plus [0, l-1] when periodic (HamiltonianModule.py:151-163; float array)."""
    pos = [[n, n + 1] for n in range(l - 1)]
    if bound_cond == 'periodic':
        pos.append([0, l - 1])
    return np.array(pos, dtype=float).reshape(-1, 2)


def positions_jigsaw_1d(length, bound_cond='open'):
    """sawtooth chain: NN bonds then next-nearest bonds between even sites (HamiltonianModule.py:166-200)."""
    if bound_cond == 'open':
        if length % 2 == 0:
            length += 1
        n_half = round((length - 1) / 2)
    else:
        if length % 2 == 1:
            length += 1
        n_half = round(length / 2)
    pos = [[n, n + 1] for n in range(length - 1)]
    if bound_cond == 'periodic':
        pos.append([0, length - 1])
        pos += [[2 * n, 2 * (n + 1)] for n in range(n_half - 1)]
        pos.append([0, length - 2])
    else:
        pos += [[2 * n, 2 * (n + 1)] for n in range(n_half)]
    return np.array(pos, dtype=float).reshape(-1, 2)


def positions_nearest_neighbor_square(width, height, bound_cond='open'):
    """site = row*width + col; horizontal bonds row by row, then vertical bonds column by column, then the periodic
    wrap-around bonds (HamiltonianModule.py:203-227)."""
    pos = []
    for r in range(height):
        pos += [[r * width + c, r * width + c + 1] for c in range(width - 1)]
    for c in range(width):
        pos += [[r * width + c, (r + 1) * width + c] for r in range(height - 1)]
    if bound_cond == 'periodic':
        pos += [[r * width, (r + 1) * width - 1] for r in range(height)]
        pos += [[c, (height - 1) * width + c] for c in range(width)]
        return np.array(pos, dtype=int).reshape(-1, 2)
    return np.array(pos, dtype=float).reshape(-1, 2)


def positions_fully_connected(n_site):
    return np.array([[i, j] for i in range(n_site - 1) for j in range(i + 1, n_site)], dtype=int).reshape(-1, 2)


def interactions_position2full_index_heisenberg_two_body(index_pos):
    """three rows [site1, site2, op1, op2] per bond: (su, sd), (sd, su), (sz, sz) with the operator numbering
    [id, sx, sy, sz, su, sd] (HamiltonianModule.py:242-261)."""
    index_pos = np.asarray(index_pos)
    rows = []
    for n in range(index_pos.shape[0]):
        i, j = int(index_pos[n, 0]), int(index_pos[n, 1])
        rows += [[i, j, 4, 5], [i, j, 5, 4], [i, j, 3, 3]]
    return np.array(rows, dtype=int).reshape(-1, 4)


def from_spin2phys_dim(spin):
    return {'half': 2, 'one': 3}.get(spin)


def hamiltonian_heisenberg(spin, jx, jy, jz, hx, hz):
    """two-site Hamiltonian jx SxSx + jy SySy + jz SzSz + hx (Sx1 + Sx2) + hz (Sz1 + Sz2) (HamiltonianModule.py:67-72)"""
    op = spin_operators(spin)
    h = jx * np.kron(op['sx'], op['sx']) + jy * np.kron(op['sy'], op['sy']).real + jz * np.kron(op['sz'], op['sz'])
    h = h + hx * (np.kron(op['id'], op['sx']) + np.kron(op['sx'], op['id']))
    h = h + hz * (np.kron(op['id'], op['sz']) + np.kron(op['sz'], op['id']))
    return h


def interactions_full_connection_two_body(l):
    """([first_site, second_site] for every pair n1 < n2, number of pairs) (HamiltonianModule.py:137-148)"""
    pairs = np.array([[n1, n2] for n1 in range(l) for n2 in range(n1 + 1, l)], dtype=float).reshape(-1, 2)
    return pairs, float(pairs.shape[0])


def hamiltonian_heisenberg_library(spin, jx, jy, jz, hx, hz):
    """the library generation's two-site Hamiltonian (library/HamiltonianModule.py:141-147): the fields enter with a MINUS sign,
    jx SxSx + jy SySy + jz SzSz - hx (Sx1 + Sx2) - hz (Sz1 + Sz2).  Used by the iDMRG / TEBD drivers of that generation."""
    return hamiltonian_heisenberg(spin, jx, jy, jz, -hx, -hz)


def hamiltonian_indexes(model, parameters):
    """rows [op1, op2, coupling] of a two-site Hamiltonian in the operator order I, sx, sy, sz, su, sd
    (library/HamiltonianModule.py:97-135).  'heisenberg': parameters = (j_ud, j_zz, h_x, h_z); 'q-ising': (j_zz, h_x)."""
    if model == 'heisenberg':
        j, jz, hx, hz = parameters
        return np.array([[4, 5, j], [5, 4, j], [3, 3, jz], [0, 1, hx], [1, 0, hx], [0, 3, hz], [3, 0, hz]], dtype=float)
    if model == 'q-ising':
        jz, hx = parameters
        return np.array([[3, 3, jz], [0, 1, hx], [1, 0, hx]], dtype=float)
    raise ValueError('hamiltonian_indexes: unknown model %r' % (model,))
