"""Term algebra and environment bookkeeping for the one-site effective Hamiltonian.

Replaces the reference's operator cache (a2-a4: all_environments_optimized / classify_and_update_env /
get_effective_operators_* / update_all_effective_id, MPSClass.py:203-325, 474-530, 633-733), which keeps one
(chi,chi) matrix per coupling term per bond and re-runs L-1 identity transfers on every local update.

Here every bond keeps *summed complementary operators* (SURVEY.md section 7.6):
  left block of bond p  (depends on sites < p):   HL[p]            sum of every term wholly left of the bond
                                                  OL[p][(i, s)]    op s of site i < p that still has a partner >= p
  right block of bond p (depends on sites >= p):  HR[p], OR[p][(j, s)]
and a bond move is ONE batched tn_env_update call.  The groups handed to the matvec are exactly the reference's
opt_env groups ('1_0_0', '0_0_1', '0_s_0', '1_s_0', '0_s_1', '1_0_1'), so results agree term by term; only the
floating-point summation order differs.  Validity is tracked per site tensor, so a gauge change made by any method
(including calculate_entanglement_spectrum, whose stale cache makes the reference's corr_* wrong, SURVEY.md
section 4) invalidates exactly the blocks that depend on it.
"""
import numpy as np


class TermTable:
    """Filtered coupling terms.  Mirrors the filters of all_environments_optimized (MPSClass.py:643-652):
    one-body terms need |c| > tol and ||op|| > tol, two-body terms |c| > tol; tol is the eigensolver tolerance."""

    def __init__(self, index1, index2, coeff1, coeff2, operators, tol):
        self.ops = [np.real(np.asarray(o)).astype(float) if np.abs(np.imag(np.asarray(o))).max() == 0 else None
                    for o in operators]
        index1 = np.asarray(index1, dtype=int).reshape(-1, 2)
        index2 = np.asarray(index2, dtype=int).reshape(-1, 4)
        coeff1 = np.asarray(coeff1, dtype=float).reshape(-1)
        coeff2 = np.asarray(coeff2, dtype=float).reshape(-1)
        self.one = []  # (site, op, c)
        for n in range(index1.shape[0]):
            site, sn, c = int(index1[n, 0]), int(index1[n, 1]), float(coeff1[n])
            if abs(c) > tol and np.linalg.norm(np.asarray(operators[sn])) > tol:
                self.one.append((site, sn, c))
        self.two = []  # (i, j, op_i, op_j, c) with i < j
        for n in range(index2.shape[0]):
            i, j, si, sj = (int(x) for x in index2[n])
            c = float(coeff2[n])
            if abs(c) > tol:
                if not i < j:
                    raise ValueError('two-body term %d has site1 >= site2 (%d, %d); the reference path is only '
                                     'correct for site1 < site2 (MPSClass.py:652,717-728)' % (n, i, j))
                self.two.append((i, j, si, sj, c))
        for _, sn, _ in self.one:
            self._real_op(sn)
        for _, _, si, sj, _ in self.two:
            self._real_op(si), self._real_op(sj)
        self.length = 1 + max([t[0] for t in self.one] + [t[1] for t in self.two] + [0])

    def _real_op(self, sn):
        if self.ops[sn] is None:
            raise ValueError('operator %d is complex; the is_real path only supports real site operators' % sn)
        return self.ops[sn]

    def site_field(self, p, d):
        """sum of the one-body operators acting on site p (the '0_s_0' group), or None"""
        m = None
        for site, sn, c in self.one:
            if site == p:
                m = c * self.ops[sn] if m is None else m + c * self.ops[sn]
        return m

    def open_left(self, p):
        """(i, s) pairs with i < p that have a partner j >= p  (operators alive on the left block of bond p)"""
        return sorted({(i, si) for i, j, si, sj, c in self.two if i < p <= j})

    def open_right(self, p):
        """(j, s) pairs with j >= p that have a partner i < p"""
        return sorted({(j, sj) for i, j, si, sj, c in self.two if i < p <= j})

    def any_left_of(self, p):
        """is there any term wholly inside sites < p"""
        return any(site < p for site, _, _ in self.one) or any(j < p for _, j, _, _, _ in self.two)

    def any_right_of(self, p):
        """is there any term wholly inside sites >= p"""
        return any(site >= p for site, _, _ in self.one) or any(i >= p for i, _, _, _, _ in self.two)

    def closing_left(self, p):
        """terms (i, p): {s_p: [(c, (i, s_i)), ...]} -- the '1_s_0' groups at site p"""
        g = {}
        for i, j, si, sj, c in self.two:
            if j == p:
                g.setdefault(sj, []).append((c, (i, si)))
        return g

    def closing_right(self, p):
        """terms (p, j): {s_p: [(c, (j, s_j)), ...]} -- the '0_s_1' groups at site p"""
        g = {}
        for i, j, si, sj, c in self.two:
            if i == p:
                g.setdefault(si, []).append((c, (j, sj)))
        return g

    def crossing(self, p):
        """terms (i, j) with i < p < j: [(c, (i, s_i), (j, s_j))] -- the '1_0_1' list at site p"""
        return [(c, (i, si), (j, sj)) for i, j, si, sj, c in self.two if i < p < j]

    def counts(self, p):
        """(K_L, K_R, n_x) of SURVEY.md section 8d for site p"""
        kl = len(self.closing_left(p)) + (1 if self.any_left_of(p) else 0)
        kr = len(self.closing_right(p)) + (1 if self.any_right_of(p + 1) else 0)
        return kl, kr, len(self.crossing(p))


class EnvCache:
    """Left/right blocks of every bond for one TermTable, kept on the device."""
    shard_min_dim = 64   # bond moves of smaller blocks are computed by every rank (a collective costs more than the GEMMs)

    def __init__(self, be, terms, length):
        self.be, self.terms, self.L = be, terms, length
        self.left = [None] * (length + 1)   # bond p: {'H': tensor|None, 'O': {(i,s): tensor}}
        self.right = [None] * (length + 1)
        self.left[0] = {'H': None, 'O': {}}
        self.right[length] = {'H': None, 'O': {}}
        self.lv = 0        # left blocks of bonds 0..lv are valid
        self.rv = length   # right blocks of bonds rv..L are valid
        self.n_bond_moves = 0
        self.merge_crossing = True   # False: one GEMM link per crossing term, the reference's literal '1_0_1' list
        # operator pairs with ops[s'] == ops[s]^T (S- = (S+)^T): the block of (i, s') is the transpose of the block of (i, s)
        # -- T^T (E (x) op^T) T = (T^T (E^T (x) op) T)^T and E^T is the partner's block by induction -- so only one of the two
        # is transferred through a bond move and the other is a transposed copy
        self.transpose_of = {}
        ops_ = terms.ops
        for s2 in range(len(ops_)):
            for s1 in range(s2):
                if ops_[s1] is not None and ops_[s2] is not None and ops_[s1].shape == ops_[s2].shape and \
                        not np.array_equal(ops_[s1], ops_[s1].T) and np.array_equal(ops_[s2], ops_[s1].T):
                    self.transpose_of.setdefault(s2, s1)
        self.comm = None             # backend communicator when the outgoing operators of a bond move are sharded over ranks

    def invalidate_site(self, n):
        """tensor n changed: left blocks of bonds > n and right blocks of bonds <= n are stale"""
        if n < self.lv:
            for p in range(n + 1, self.lv + 1):
                self.left[p] = None
            self.lv = n
        if n + 1 > self.rv:
            for p in range(self.rv, n + 1):
                self.right[p] = None
            self.rv = n + 1

    def invalidate_all(self):
        for n in range(self.L):
            self.invalidate_site(n)

    # ---- bond moves ----
    def _env_update(self, direction, T, outputs):
        """tn_env_update for every outgoing operator of a bond; with several ranks the operators are dealt round-robin
        (heaviest first), every rank writes its own into its rows of one (world * k, e, e) buffer and ONE in-place
        all-gather completes it on all ranks (one writer per operator: bit-identical blocks everywhere; the blocks
        handed back are views of that buffer)"""
        comm = self.comm
        e_dim = T.shape[2] if direction == 0 else T.shape[0]
        if comm is None or comm.world == 1 or len(outputs) < 2 or e_dim < self.shard_min_dim:
            return self.be.env_update(direction, T, outputs)
        rank, world = comm.rank, comm.world
        order = sorted(range(len(outputs)), key=lambda j: -len(outputs[j]))
        per = (len(outputs) + world - 1) // world
        buf = self.be.empty(world * per, e_dim, e_dim)
        place = {j: (i % world) * per + i // world for i, j in enumerate(order)}
        mine = [j for i, j in enumerate(order) if i % world == rank]
        if mine:
            self.be.env_update(direction, T, [outputs[j] for j in mine], outs=[buf[place[j]] for j in mine])
        comm.allgather_inplace(buf)
        return [buf[place[j]] for j in range(len(outputs))]

    def _mirrored(self, live):
        """{(site, s'): (site, s)} for the live operators whose block is the transpose of another live block"""
        have = set(live)
        return {(i, s2): (i, self.transpose_of[s2]) for (i, s2) in live
                if s2 in self.transpose_of and (i, self.transpose_of[s2]) in have}

    def _lincomb(self, block, pairs):
        xs = [block['O'][key] for _, key in pairs]
        cs = [c for c, _ in pairs]
        if len(xs) == 1 and cs[0] == 1.0:
            return xs[0]
        return self.be.lincomb(xs, cs)

    def _advance_left(self, p, T):
        """left block of bond p+1 from bond p and the (left-orthonormal) tensor of site p"""
        t, cur = self.terms, self.left[p]
        outputs, keys = [], []
        links = []
        if cur['H'] is not None:
            links.append((cur['H'], None))
        for s, pairs in sorted(t.closing_left(p).items()):
            links.append((self._lincomb(cur, pairs), t.ops[s]))
        field = t.site_field(p, T.shape[1])
        if field is not None:
            links.append((None, field))
        if links:
            outputs.append(links)
            keys.append('H')
        live = t.open_left(p + 1)
        mirrored = self._mirrored(live)
        for key in live:
            if key in mirrored:
                continue
            i, s = key
            outputs.append([(None, t.ops[s])] if i == p else [(cur['O'][key], None)])
            keys.append(key)
        new = {'H': None, 'O': {}}
        if outputs:
            res = self._env_update(0, T, outputs)
            for key, mat in zip(keys, res):
                if key == 'H':
                    new['H'] = mat
                else:
                    new['O'][key] = mat
        for key, partner in mirrored.items():
            new['O'][key] = new['O'][partner].t().contiguous()
        self.left[p + 1] = new
        self.n_bond_moves += 1

    def _advance_right(self, p, T):
        """right block of bond p from bond p+1 and the (right-orthonormal) tensor of site p"""
        t, cur = self.terms, self.right[p + 1]
        outputs, keys = [], []
        links = []
        if cur['H'] is not None:
            links.append((cur['H'], None))
        for s, pairs in sorted(t.closing_right(p).items()):
            links.append((self._lincomb(cur, pairs), t.ops[s]))
        field = t.site_field(p, T.shape[1])
        if field is not None:
            links.append((None, field))
        if links:
            outputs.append(links)
            keys.append('H')
        live = t.open_right(p)
        mirrored = self._mirrored(live)
        for key in live:
            if key in mirrored:
                continue
            j, s = key
            outputs.append([(None, t.ops[s])] if j == p else [(cur['O'][key], None)])
            keys.append(key)
        new = {'H': None, 'O': {}}
        if outputs:
            res = self._env_update(1, T, outputs)
            for key, mat in zip(keys, res):
                if key == 'H':
                    new['H'] = mat
                else:
                    new['O'][key] = mat
        for key, partner in mirrored.items():
            new['O'][key] = new['O'][partner].t().contiguous()
        self.right[p] = new
        self.n_bond_moves += 1

    def ensure(self, p, mps, width=1):
        """make the left block of bond p and the right block of bond p+width valid (centre at p; width = 2 for the
        two-site update)"""
        while self.lv < p:
            self._advance_left(self.lv, mps[self.lv])
            self.lv += 1
        while self.rv > p + width:
            self._advance_right(self.rv - 1, mps[self.rv - 1])
            self.rv -= 1

    # ---- the opt_env groups of site p (MPSClass.py:684-733) ----
    def groups(self, p, d):
        t, lb, rb = self.terms, self.left[p], self.right[p + 1]
        g = {'HL': lb['H'], 'HR': rb['H'], 'M': t.site_field(p, d), 'LS': [], 'ls_ops': [], 'RS': [], 'rs_ops': [],
             'XL': [], 'XR': [], 'x_coeff': []}
        for s, pairs in sorted(t.closing_left(p).items()):
            g['LS'].append(self._lincomb(lb, pairs))
            g['ls_ops'].append(t.ops[s])
        for s, pairs in sorted(t.closing_right(p).items()):
            g['RS'].append(self._lincomb(rb, pairs))
            g['rs_ops'].append(t.ops[s])
        # crossing terms ('1_0_1'): sum_i c_i OL[lk_i] (x) OR[rk_i].  Terms sharing an operator on one side are merged
        # first, sum_i c_i L_k (x) R_i = L_k (x) (sum_i c_i R_i), exactly as the reference merges the partners of a
        # '1_s_0' key (MPSClass.py:717-728); the side with fewer distinct operators keeps its matrices, the other side
        # becomes one tn_lincomb per group.  Same operator, fewer GEMM links (42 -> 21 at the widest 6x6 J1-J2 site).
        cross = t.crossing(p)
        by_left, by_right = {}, {}
        for c, lk, rk in cross:
            by_left.setdefault(lk, []).append((c, rk))
            by_right.setdefault(rk, []).append((c, lk))
        if not self.merge_crossing:
            for c, lk, rk in cross:
                g['XL'].append(lb['O'][lk])
                g['XR'].append(rb['O'][rk])
                g['x_coeff'].append(c)
        elif len(by_left) <= len(by_right):
            for lk in sorted(by_left):
                g['XL'].append(lb['O'][lk])
                g['XR'].append(self._lincomb(rb, by_left[lk]))
                g['x_coeff'].append(1.0)
        else:
            for rk in sorted(by_right):
                g['XL'].append(self._lincomb(lb, by_right[rk]))
                g['XR'].append(rb['O'][rk])
                g['x_coeff'].append(1.0)
        g['n_x_reference'] = len(cross)
        return g

    # ---- two-site effective Hamiltonian of sites (p, p+1) ----
    def groups_two_site(self, p, d):
        """opt_env-style groups for the pair (p, p+1) acting on theta[a, (s1 s2), b] with the combined physical index
        s = s1*d + s2 (dimension d*d): same six kinds of groups as the one-site case (MPSClass.py:684-733), with site
        operators op (x) 1 / 1 (x) op and the in-pair terms folded into the on-site matrix M.  This is the term-summed
        two-site matvec of the reference's White-style iDMRG (library/MPSClass.py:1688-1707) for an arbitrary term list."""
        t, lb, rb = self.terms, self.left[p], self.right[p + 2]
        P, Q, eye = p, p + 1, np.eye(d)
        M = np.zeros((d * d, d * d))
        has_m = False
        for site, sn, c in t.one:
            if site == P:
                M += c * np.kron(t.ops[sn], eye)
                has_m = True
            elif site == Q:
                M += c * np.kron(eye, t.ops[sn])
                has_m = True
        ls, rs, cross = {}, {}, []
        for i, j, si, sj, c in t.two:
            if i == P and j == Q:
                M += c * np.kron(t.ops[si], t.ops[sj])
                has_m = True
            elif j == P:
                ls.setdefault((0, sj), []).append((c, (i, si)))
            elif j == Q and i < P:
                ls.setdefault((1, sj), []).append((c, (i, si)))
            elif i == P and j > Q:
                rs.setdefault((0, si), []).append((c, (j, sj)))
            elif i == Q and j > Q:
                rs.setdefault((1, si), []).append((c, (j, sj)))
            elif i < P and j > Q:
                cross.append((c, (i, si), (j, sj)))

        def pair_op(slot, sn):
            return np.kron(t.ops[sn], eye) if slot == 0 else np.kron(eye, t.ops[sn])

        g = {'HL': lb['H'], 'HR': rb['H'], 'M': M if has_m else None, 'LS': [], 'ls_ops': [], 'RS': [], 'rs_ops': [],
             'XL': [], 'XR': [], 'x_coeff': []}
        for (slot, sn), pairs in sorted(ls.items()):
            g['LS'].append(self._lincomb(lb, pairs))
            g['ls_ops'].append(pair_op(slot, sn))
        for (slot, sn), pairs in sorted(rs.items()):
            g['RS'].append(self._lincomb(rb, pairs))
            g['rs_ops'].append(pair_op(slot, sn))
        by_left, by_right = {}, {}
        for c, lk, rk in cross:
            by_left.setdefault(lk, []).append((c, rk))
            by_right.setdefault(rk, []).append((c, lk))
        if len(by_left) <= len(by_right):
            for lk in sorted(by_left):
                g['XL'].append(lb['O'][lk])
                g['XR'].append(self._lincomb(rb, by_left[lk]))
                g['x_coeff'].append(1.0)
        else:
            for rk in sorted(by_right):
                g['XL'].append(self._lincomb(lb, by_right[rk]))
                g['XR'].append(rb['O'][rk])
                g['x_coeff'].append(1.0)
        g['n_x_reference'] = len(cross)
        return g

    def plan_two_site(self, p, mps, rank=0, world=1, rows=None):
        self.ensure(p, mps, width=2)
        a, d, _ = mps[p].shape
        b = mps[p + 1].shape[2]
        g = self.groups_two_site(p, d)
        plan = self.be.effh_plan((a, d * d, b), g['HL'], g['HR'], g['M'], g['LS'], g['ls_ops'], g['RS'], g['rs_ops'],
                                 g['XL'], g['XR'], g['x_coeff'], rank=rank, world=world, **({'rows': rows} if rows is not None else {}))
        kl = (1 if g['HL'] is not None else 0) + len(g['LS'])
        kr = (1 if g['HR'] is not None else 0) + len(g['RS'])
        plan.flops_algorithmic = 2.0 * a * d * d * b * (a * (kl + g['n_x_reference']) + b * (kr + g['n_x_reference']))
        return plan

    def plan(self, p, mps, rank=0, world=1, rows=None):
        self.ensure(p, mps)
        a, d, b = mps[p].shape
        g = self.groups(p, d)
        plan = self.be.effh_plan((a, d, b), g['HL'], g['HR'], g['M'], g['LS'], g['ls_ops'], g['RS'], g['rs_ops'],
                                 g['XL'], g['XR'], g['x_coeff'], rank=rank, world=world, **({'rows': rows} if rows is not None else {}))
        # algorithmic flop of one matvec = the reference's own grouping, whatever the kernel executes
        kl = (1 if g['HL'] is not None else 0) + len(g['LS'])
        kr = (1 if g['HR'] is not None else 0) + len(g['RS'])
        plan.flops_algorithmic = 2.0 * a * d * b * (a * (kl + g['n_x_reference']) + b * (kr + g['n_x_reference']))
        return plan


def expect_products(be, mps, center, ops, terms, comm=None, min_terms=16):
    """expectation values of operator products (see _expect_products_local).  With a communicator (bit-identical replicas on
    several GPUs) the terms are dealt to the ranks by their leading site, every rank contracts its share, and one all-reduce of the value vector (each entry has exactly one
    non-zero contribution, so the sum is exact and identical everywhere) completes the result on all ranks."""
    if comm is None or comm.world == 1 or len(terms) < min_terms:
        return _expect_products_local(be, mps, center, ops, terms)
    # contiguous blocks of leading sites, balanced by the number of bonds the terms' environments live on: terms that share
    # a prefix environment (same leading site) stay together and the closing blocks a rank needs cluster around its sites
    L = len(mps)
    weight = np.zeros(L)
    for term in terms:
        weight[int(term[0][0])] += 1 + int(term[-1][0]) - int(term[0][0])
    cum = np.cumsum(weight)
    owner = [min(comm.world - 1, int((cum[s] - weight[s] / 2) * comm.world / cum[-1])) for s in range(L)]
    mine = [i for i, term in enumerate(terms) if owner[int(term[0][0])] == comm.rank]
    vals = np.zeros(len(terms))
    if mine:
        vals[mine] = _expect_products_local(be, mps, center, ops, [terms[i] for i in mine])
    buf = be.from_numpy(vals)
    comm.allreduce(buf)
    return be.to_numpy(buf)


def _expect_products_local(be, mps, center, ops, terms):
    """<prod_k op[s_k](site_k)> for every term in `terms` (list of tuples of (site, op_id), sites strictly increasing
    inside a term) on a centre-orthogonal MPS (sites < center left-, sites > center right-orthonormal).
    a10: observation_s1 / observation_s1_s2 (MPSClass.py:857-909) restated as one left-to-right pass:
      * operator-carrying environments are shared by all terms with the same leading (site, op); the far sides are the
        identity or a 'density' chain rho from the centre; every bond is one batched tn_env_update call;
      * a term is closed on its last site by the inner product of its environment with the block R(site, op) -- the last
        operator contracted with the site tensor and whatever lies to its right -- built once per distinct (site, op), so closing
        costs a dot product instead of a transfer per term;
      * the state is real, so <A> = <A^T>: a term whose operators are the transposes of another term's (S- S+ vs S+ S-) is not
        contracted twice.
    Returns a float numpy array of len(terms)."""
    L = len(mps)
    if not terms:
        return np.zeros(0)
    terms_in = [tuple((int(s), int(o)) for s, o in term) for term in terms]
    for term in terms_in:
        sites = [s for s, _ in term]
        if sites != sorted(set(sites)) or not (0 <= sites[0] and sites[-1] < L):
            raise ValueError('observable sites must be strictly increasing and inside the chain: %r' % (term,))
    # transpose partner of every operator that occurs (None: its transpose is not in the operator list)
    used = sorted({o for term in terms_in for _, o in term})
    partner = {}
    for o in used:
        partner[o] = next((o2 for o2 in range(len(ops)) if ops[o2] is not None and ops[o2].shape == ops[o].shape
                           and np.array_equal(ops[o2], ops[o].T)), None)
    canon = []
    for term in terms_in:
        if all(partner[o] is not None for _, o in term):
            canon.append(min(term, tuple((s, partner[o]) for s, o in term)))
        else:
            canon.append(term)
    terms = sorted(set(canon))
    first = min(term[0][0] for term in terms)
    last = max(term[-1][0] for term in terms)
    # right density chain: rhoR[q] closes bond q when every site >= q carries no operator and q <= center
    rhoR = {}
    lo = min(term[-1][0] for term in terms) + 1  # smallest closing bond
    if lo <= center:
        cur = None
        for q in range(center, lo - 1, -1):
            cur = be.env_update(1, mps[q], [[(cur, None)]])[0]
            rhoR[q] = cur
    slots, slot_of = [], {}
    # active[(prefix term tuple)] = environment at the current bond carrying those operators; () is the plain density
    active = {}
    need_rho_left = any(term[0][0] > center for term in terms)
    for q in range(min(first, center) if need_rho_left else first, last + 1):
        # ---- close the terms that end on this site against R(q, op) ----
        closing = [term for term in terms if term[-1][0] == q]
        if closing:
            close_ops = sorted({term[-1][1] for term in closing})
            right = rhoR[q + 1] if q < center else None
            blocks = dict(zip(close_ops, be.env_update(1, mps[q], [[(right, ops[o])] for o in close_ops])))
            for term in closing:
                parent = term[:-1]
                if parent == ():
                    env = active.get(()) if q > center else None      # identity unless a density has left the centre
                    if q > center and env is None:
                        raise RuntimeError('internal: missing density chain at site %d' % q)
                else:
                    env = active[parent]
                blk = blocks[term[-1][1]]
                slot_of[term] = len(slots)
                slots.append(be.trace(blk) if env is None else be.dot(env, blk))
                if len(slots) >= 2048:
                    raise RuntimeError('too many observables in one call')
        # ---- which prefixes must exist on bond q+1 ----
        outputs, keys = [], []
        want = set()
        for term in terms:
            sites = [s for s, _ in term]
            if sites[0] > q:
                if q >= center and sites[0] > center:
                    want.add(())      # density still travelling towards the first operator
                continue
            if sites[-1] <= q:
                continue
            k = sum(1 for s in sites if s <= q)  # operators absorbed up to and including site q
            want.add(term[:k])
        for pre in sorted(want):
            if pre and pre[-1][0] == q:      # absorbs an operator on this site
                parent = pre[:-1]
                op = ops[pre[-1][1]]
            else:
                parent, op = pre, None
            if parent == ():
                env = active.get(()) if q > center else None   # identity unless a density has left the centre
                if q > center and env is None:
                    raise RuntimeError('internal: missing density chain at site %d' % q)
            else:
                env = active[parent]
            if env is None and op is None:
                # identity through a left-orthonormal site stays the identity; only q == center starts the density
                if q < center:
                    continue
            outputs.append([(env, op)])
            keys.append(pre)
        new_active = {}
        if outputs:
            res = be.env_update(0, mps[q], outputs)
            new_active = dict(zip(keys, res))
        active = new_active
    vals = be.scalars_to_host(slots)
    return np.array([vals[slot_of[c]] for c in canon])
