"""
Tolerance-process expectations conditional on a primary trajectory:
the dense reference functions of raoteh/sampler/_tmjp_dense.py that sit on the hot
path of the Rao-Blackwellised tolerance sampler (examples/p53/blink.py:54-67).

Given a primary trajectory, every tolerance class is an independent piecewise
homogeneous 3-state MJP (off, on, absorbing) on the same tree
(`get_inhomogeneous_mjp`, :965).  Its expectations are the generic hot path with
S = 3 and one rate matrix per edge: rt_expm_batched, rt_support_sets,
rt_prune_loglik, rt_posterior_stats, rt_frechet_contract replace
pyfelscore.get_tolerance_rate_matrix (:239), the `kitchen_sink` calls (:297) and
pyfelscore.get_tolerance_expectations (:339).
"""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np
from scipy import special

from . import _core, _util
from ..lowering import TreeSchedule, check_square_dense

__all__ = []


def get_two_state_tolerance_distn(rate_off, rate_on):
    """raoteh/sampler/_tmjp_dense.py:352-377"""
    if (rate_off < 0) or (rate_on < 0):
        raise ValueError('rates must be non-negative')
    total = rate_off + rate_on
    if total <= 0:
        raise ValueError('the total tolerance rate must be positive')
    return np.array([rate_off, rate_on], dtype=float) / total


def get_three_state_tolerance_distn(rate_off, rate_on):
    """raoteh/sampler/_tmjp_dense.py:380-404"""
    if (rate_off < 0) or (rate_on < 0):
        raise ValueError('rates must be non-negative')
    total = rate_off + rate_on
    if total <= 0:
        raise ValueError('the total tolerance rate must be positive')
    return np.array([rate_off, rate_on, 0], dtype=float) / total


def get_primary_state_to_absorption_rate(Q_primary, primary_to_part, tolerance_class):
    """raoteh/sampler/_tmjp_dense.py:931-962"""
    check_square_dense(Q_primary)
    nprimary = len(primary_to_part)
    out = {}
    for sa in range(nprimary):
        out[sa] = sum(Q_primary[sa, sb] for sb in range(nprimary)
                      if sb != sa and primary_to_part[sb] == tolerance_class)
    return out


def get_inhomogeneous_mjp(primary_to_part, rate_on, rate_off, Q_primary, T_primary, root,
                          T_primary_edges, tolerance_class):
    """raoteh/sampler/_tmjp_dense.py:965-1085 (spec raoteh/sampler/_tmjp.py:815-902):
    per edge of the primary trajectory, Q_tol = [[-on, on, 0], [off', -off'-abs, abs],
    [0, 0, 0]] with off' = 0 and state 'off' forbidden at both ends where the primary
    state belongs to the class.  Returns (T_tol, node_to_allowed_tolerances)."""
    check_square_dense(Q_primary)
    absorb = get_primary_state_to_absorption_rate(Q_primary, primary_to_part, tolerance_class)
    T_tol = nx.Graph()
    allowed = dict((n, {0, 1}) for n in T_primary)
    for na, nb in T_primary_edges:
        s = T_primary[na][nb]['state']
        same = primary_to_part[s] == tolerance_class
        off = 0.0 if same else rate_off
        r = absorb[s]
        Q = np.array([[-rate_on, rate_on, 0.0], [off, -off - r, r], [0.0, 0.0, 0.0]])
        T_tol.add_edge(na, nb, weight=T_primary[na][nb]['weight'], Q=Q)
        if same:
            allowed[na].discard(0)
            allowed[nb].discard(0)
    return T_tol, allowed


def get_expected_tolerance_history_statistics(T, node_to_allowed_states, root, root_distn=None):
    """raoteh/sampler/_tmjp_dense.py:246-349 -> (dwell[2], root posterior[3],
    transitions[2,2], absorption expectation)."""
    if root not in T:
        raise ValueError('the specified root is not in the tree')
    sched = TreeSchedule.from_nx(T, root)
    P, Qs = _core.expm_edges(sched, T, 3, None)
    ev = _core.Evaluation(sched, P, root_distn, 3)
    mask = _core.mask_from_allowed(sched, node_to_allowed_states, 3)
    mask = ev.support(mask, passes=3)
    ll, status, pmap = ev.upward_masks(mask)
    if status != 0:
        raise _util.NumericalZeroProb('the denominator is zero')
    D, J = ev.downward()
    M = ev.expectations(Qs, sched.length)
    dwell = np.zeros(2)
    trans = np.zeros((2, 2))
    absorption = 0.0
    for i in range(1, sched.n):
        Q = Qs[i]
        dwell[0] += M[i, 0, 0]
        dwell[1] += M[i, 1, 1]
        trans[0, 1] += Q[0, 1] * M[i, 0, 1]
        trans[1, 0] += Q[1, 0] * M[i, 1, 0]
        absorption += Q[1, 2] * M[i, 1, 1]
    return dwell, D[0], trans, absorption


def get_tolerance_summary(primary_to_part, rate_on, rate_off, Q_primary, T_primary, root,
                          disease_data=None):
    """raoteh/sampler/_tmjp_dense.py:724-855 -> the seven tolerance expectations."""
    total_weight = T_primary.size(weight='weight')
    nparts = len(set(primary_to_part.values()))
    tolerance_distn = get_three_state_tolerance_distn(rate_off, rate_on)
    edges = list(nx.bfs_edges(T_primary, root))
    ngains = nlosses = dwell_on = initial_on = nabsorptions = 0.0
    for tolerance_class in range(nparts):
        T_tol, allowed = get_inhomogeneous_mjp(primary_to_part, rate_on, rate_off, Q_primary,
                                               T_primary, root, edges, tolerance_class)
        if disease_data is not None:
            for node, tol_set in disease_data[tolerance_class].items():
                allowed[node].intersection_update(tol_set)
        dwell, post_root, trans, absorb = get_expected_tolerance_history_statistics(
            T_tol, allowed, root, root_distn=tolerance_distn)
        dwell_on += dwell[1]
        ngains += trans[0, 1]
        nlosses += trans[1, 0]
        initial_on += post_root[1]
        nabsorptions += absorb
    initial_off = nparts - initial_on
    dwell_off = total_weight * nparts - dwell_on
    return (initial_on, initial_off, dwell_on, dwell_off, nabsorptions, ngains, nlosses)


def get_tolerance_ll_contribs(rate_on, rate_off, total_tree_length,
                              expected_initial_on, expected_initial_off,
                              expected_dwell_on, expected_dwell_off,
                              expected_nabsorptions, expected_ngains, expected_nlosses):
    """raoteh/sampler/_tmjp_dense.py:858-928"""
    tolerance_distn = get_three_state_tolerance_distn(rate_off, rate_on)
    init_ll = (special.xlogy(expected_initial_on - 1, tolerance_distn[1]) +
               special.xlogy(expected_initial_off, tolerance_distn[0]))
    dwell_prim = -expected_nabsorptions
    dwell_tol = -(expected_dwell_off * rate_on + (expected_dwell_on - total_tree_length) * rate_off)
    trans_ll = special.xlogy(expected_ngains, rate_on) + special.xlogy(expected_nlosses, rate_off)
    return init_ll, dwell_prim, dwell_tol, trans_ll


def get_tolerance_process_log_likelihood(Q_primary, primary_to_part, T_primary,
                                         rate_off, rate_on, primary_root_distn, root):
    """raoteh/sampler/_tmjp_dense.py:407-505: log-likelihood of a primary trajectory under the
    compound process with the tolerance histories integrated out -- log prior of the root state,
    xlogy(count, rate) over the primary transitions, and per tolerance class the log-likelihood of
    its piecewise homogeneous 3-state process (rt_expm_batched + rt_prune_loglik with one rate
    matrix per edge).  The batched device form is column 7 of the tolerance summary
    (raoteh_b200.tmjp.ToleranceChains.tolerance_log_likelihood)."""
    from . import _mjp_dense
    if root is None:
        raise ValueError('unspecified root')
    if root not in T_primary:
        raise ValueError('the specified root is not a node in the tree')
    check_square_dense(Q_primary)
    nprimary = len(primary_to_part)
    tolerance_distn = get_three_state_tolerance_distn(rate_off, rate_on)
    root_state, transitions = _mjp_dense.get_history_root_state_and_transitions(
        T_primary, nprimary, root=root)
    log_likelihood = np.log(primary_root_distn[root_state])
    off = ~np.eye(nprimary, dtype=bool)
    log_likelihood += special.xlogy(transitions[off], np.asarray(Q_primary)[off]).sum()
    edges = list(nx.bfs_edges(T_primary, root))
    for tolerance_class in sorted(set(primary_to_part.values())):
        if primary_to_part[root_state] == tolerance_class:
            prior = np.array([0.0, 1.0, 0.0])
        else:
            prior = tolerance_distn
        T_tol, allowed = get_inhomogeneous_mjp(primary_to_part, rate_on, rate_off, Q_primary,
                                               T_primary, root, edges, tolerance_class)
        likelihood = _mjp_dense.get_likelihood(T_tol, allowed, root, 3, root_distn=prior, Q_default=None)
        log_likelihood += np.log(likelihood)
    return log_likelihood


class CompoundNegLL(object):
    """raoteh/sampler/_tmjp_util.py:4-25"""

    def __init__(self, init_prim, init_tol, dwell_prim, dwell_tol, trans_prim, trans_tol):
        self.init_prim, self.init_tol = init_prim, init_tol
        self.dwell_prim, self.dwell_tol = dwell_prim, dwell_tol
        self.trans_prim, self.trans_tol = trans_prim, trans_tol

    @property
    def init(self):
        return self.init_prim + self.init_tol

    @property
    def dwell(self):
        return self.dwell_prim + self.dwell_tol

    @property
    def trans(self):
        return self.trans_prim + self.trans_tol


def differential_entropy_helper(ctm, post_root_distn, post_dwell_times, post_transitions):
    """raoteh/sampler/_tmjp_dense.py:508-721 (sparse twin _tmjp.py:217-349): the negative expected
    log-likelihood of the compound trajectory from posterior expectations over the COMPOUND state
    space (1-D root posterior and dwell times, 2-D expected transition counts), separated into
    primary / tolerance parts of the initial-state, dwell and transition terms.  Vectorised over
    the compound states; the three consistency checks of the reference are kept."""
    from . import _mjp_dense
    if ctm.Q_compound is None:
        raise ValueError('call ctm.init_compound() first')
    Qc = np.asarray(ctm.Q_compound, dtype=float)
    n = ctm.ncompound
    post_root = np.asarray(post_root_distn, dtype=float)
    dwell = np.asarray(post_dwell_times, dtype=float)
    trans = np.asarray(post_transitions, dtype=float)
    init, dwl, trn = _mjp_dense.differential_entropy_helper(
        Qc, np.asarray(ctm.compound_distn, dtype=float), post_root, dwell, trans)
    prim = np.asarray(ctm.compound_to_primary)
    tols = np.asarray(ctm.compound_to_tolerances)
    part_of = np.array([ctm.primary_to_part[p] for p in range(ctm.nprimary)])
    # initial state (states with zero posterior mass are skipped, :608-652): primary prior, then
    # the tolerance classes other than the primary state's own
    sel = post_root != 0
    init_prim = -special.xlogy(post_root[sel], np.asarray(ctm.primary_distn, dtype=float)[prim[sel]]).sum()
    own_on = tols[np.arange(n), part_of[prim]] == 1
    on_count = tols.sum(axis=1)
    off_count = ctm.nparts - on_count
    t_off, t_on = ctm.tolerance_distn[0], ctm.tolerance_distn[1]
    init_tol = 0.0
    if (sel & ~own_on).any():
        init_tol = np.inf          # the reference assigns an infinite cost to such a formal state
    if (sel & (on_count > 0)).any():
        m = sel & (on_count > 0)
        init_tol = (init_tol - special.xlogy(post_root[m] * (on_count[m] - 1), t_on).sum()) if t_on else np.inf
    if (sel & (off_count > 0)).any():
        m = sel & (off_count > 0)
        init_tol = (init_tol - special.xlogy(post_root[m] * off_count[m], t_off).sum()) if t_off else np.inf
    if not np.allclose(init_prim + init_tol, init):
        raise Exception('internal differential entropy calculation error: %s + %s = %s but expected %s'
                        % (init_prim, init_tol, init_prim + init_tol, init))
    # dwell times (:667-690): tolerance part = blinking of the other classes, primary part = rates
    # of primary changes out of the compound state
    dwell_tol = float((dwell * off_count * ctm.rate_on + dwell * (on_count - 1) * ctm.rate_off).sum())
    same_prim = prim[:, None] == prim[None, :]
    absorption = np.where(~same_prim, Qc, 0.0).sum(axis=1)
    dwell_prim = float((dwell * absorption).sum())
    if not np.allclose(dwell_prim + dwell_tol, dwl):
        raise Exception('internal error')
    # transitions (:697-707): every ordered pair, split by whether the primary state changes
    x = special.xlogy(trans, Qc)
    trans_prim = -x[~same_prim].sum()
    trans_tol = -x[same_prim].sum()
    if not np.allclose(trans_prim + trans_tol, trn):
        raise Exception('internal error')
    return CompoundNegLL(init_prim, init_tol, dwell_prim, dwell_tol, trans_prim, trans_tol)


# ---------------------------------------------------------------------------
# the compound tolerance model (raoteh/sampler/_tmjp_dense.py:35-179)
# ---------------------------------------------------------------------------
class CompoundToleranceModel(object):
    """Read-only description of the compound process: dense primary rate matrix with
    diagonal, primary distribution (1d ndarray), primary state -> tolerance class, blink
    rates.  Same attributes as raoteh/sampler/_tmjp_dense.py:35-82; init_compound() builds
    the compound state space exactly as :84-179."""

    def __init__(self, Q_primary, primary_distn, primary_to_part, rate_on, rate_off):
        self.Q_primary = Q_primary
        self.primary_distn = primary_distn
        self.primary_to_part = primary_to_part
        self.rate_on = rate_on
        self.rate_off = rate_off
        self.nprimary = len(primary_to_part)
        self.nparts = len(set(primary_to_part.values()))
        self.ncompound = int(np.ldexp(self.nprimary, self.nparts))
        self.tolerance_distn = get_two_state_tolerance_distn(rate_off, rate_on)
        self.Q_compound = None
        self.compound_distn = None
        self.compound_to_primary = None
        self.compound_to_tolerances = None

    def init_compound(self):
        import itertools
        if self.Q_compound is not None:
            raise Exception('compound attributes should be initialized only once')
        if self.ncompound > 1e6:
            raise Exception('the compound state space is too big')
        self.compound_to_primary = []
        self.compound_to_tolerances = []
        for primary, tolerances in itertools.product(
                range(self.nprimary), itertools.product((0, 1), repeat=self.nparts)):
            self.compound_to_primary.append(primary)
            self.compound_to_tolerances.append(tolerances)
        prim = np.array(self.compound_to_primary)
        tols = np.array(self.compound_to_tolerances)
        part_of = np.array([self.primary_to_part[p] for p in range(self.nprimary)])
        n = self.ncompound
        own_on = tols[np.arange(n), part_of[prim]] == 1
        # P(tolerances) over the classes other than the primary state's own
        logp = np.where(tols == 1, self.tolerance_distn[1], self.tolerance_distn[0]).astype(float)
        logp[np.arange(n), part_of[prim]] = 1.0
        self.compound_distn = np.where(own_on, np.asarray(self.primary_distn)[prim] * logp.prod(axis=1), 0.0)
        for name, d in (('primary', np.asarray(self.primary_distn)), ('tolerance', self.tolerance_distn),
                        ('compound', self.compound_distn)):
            if not np.allclose(d.sum(), 1):
                raise Exception('internal error')
        Q = np.zeros((n, n), dtype=float)
        diff = tols[:, None, :] != tols[None, :, :]
        hdist = diff.sum(axis=2)
        same_prim = prim[:, None] == prim[None, :]
        # one tolerance change, same primary state, not the primary state's own class
        ii, jj = np.nonzero((hdist == 1) & same_prim)
        cls = diff[ii, jj].argmax(axis=1)
        ok = cls != part_of[prim[ii]]
        ii, jj, cls = ii[ok], jj[ok], cls[ok]
        Q[ii, jj] = np.where(tols[jj, cls] == 1, self.rate_on, self.rate_off)
        # a primary change into a tolerated class, tolerances unchanged
        ii, jj = np.nonzero((hdist == 0) & ~same_prim)
        ok = (tols[ii, part_of[prim[jj]]] == 1) & (np.asarray(self.Q_primary)[prim[ii], prim[jj]] != 0)
        ii, jj = ii[ok], jj[ok]
        Q[ii, jj] = np.asarray(self.Q_primary)[prim[ii], prim[jj]]
        Q -= np.diag(Q.sum(axis=1))
        self.Q_compound = Q


def get_tolerance_rate_matrix(rate_off, rate_on):
    """raoteh/sampler/_tmjp_dense.py:182-202"""
    return np.array([[-rate_on, rate_on], [rate_off, -rate_off]], dtype=float)


def get_primary_proposal_rate_matrix(Q_primary, primary_to_part, tolerance_distn):
    """raoteh/sampler/_tmjp_dense.py:1081-1131: between-class rates scaled by P(on)."""
    check_square_dense(Q_primary)
    nprimary = len(primary_to_part)
    part = np.array([primary_to_part[s] for s in range(nprimary)])
    Q = np.array(Q_primary, dtype=float)
    np.fill_diagonal(Q, 0.0)
    cross = part[:, None] != part[None, :]
    Q[cross] *= tolerance_distn[1]
    Q -= np.diag(Q.sum(axis=1))
    return Q
