"""
Markov jump process likelihood and closed-form posterior expectations on trees
with dense rate matrices -- the reference's `raoteh.sampler._mjp_dense`
signatures (raoteh/sampler/_mjp_dense.py), computed by the CUDA hot path:
rt_expm_batched (per-edge P), rt_support_sets + rt_prune_loglik (pruning),
rt_posterior_stats (down pass, J/P weights), rt_frechet_contract (one Frechet
derivative per edge instead of S + nnz(Q) scipy calls per edge).
"""
from __future__ import division, print_function, absolute_import

from collections import defaultdict

import networkx as nx
import numpy as np
from scipy import special

from . import _core, _util, _mc0_dense, _mcy_dense
from ..lowering import TreeSchedule, check_square_dense

__all__ = []


def get_total_rates(Q):
    """raoteh/sampler/_mjp_dense.py:28-44"""
    check_square_dense(Q)
    return -np.diag(Q)


def get_conditional_transition_matrix(Q, total_rates=None):
    """raoteh/sampler/_mjp_dense.py:47-70"""
    check_square_dense(Q)
    if total_rates is None:
        total_rates = get_total_rates(Q)
    P = Q / total_rates
    np.fill_diagonal(P, 0)
    return P


def get_history_dwell_times(T, nstates):
    """raoteh/sampler/_mjp_dense.py:73-95"""
    dwell = np.zeros(nstates, dtype=float)
    for a, b in T.edges():
        dwell[T[a][b]['state']] += T[a][b]['weight']
    return dwell


def get_history_root_state_and_transitions(T, nstates, root=None):
    """raoteh/sampler/_mjp_dense.py:98-147"""
    degrees = dict(T.degree())
    if root is None:
        root = _util.get_arbitrary_tip(T, degrees)
    root_states = [T[root][b]['state'] for b in T[root]]
    if len(set(root_states)) != 1:
        raise ValueError('the root does not have a well defined state')
    counts = np.zeros((nstates, nstates), dtype=float)
    successors = nx.dfs_successors(T, root)
    for a, b in nx.bfs_edges(T, root):
        if degrees[b] == 2:
            c = _util.get_first_element(successors[b])
            sa, sb = T[a][b]['state'], T[b][c]['state']
            if sa != sb:
                counts[sa, sb] += 1
    return root_states[0], counts


def get_history_statistics(T, nstates, root=None):
    """raoteh/sampler/_mjp_dense.py:150-184 -> (dwell times, root state, transition counts)."""
    dwell = get_history_dwell_times(T, nstates)
    root_state, transitions = get_history_root_state_and_transitions(T, nstates, root=root)
    return dwell, root_state, transitions


def get_trajectory_log_likelihood(T_aug, root, prior_root_distn, Q_default, nstates):
    """raoteh/sampler/_mjp_dense.py:188-241"""
    nstates = prior_root_distn.shape[0]
    total_rates = get_total_rates(Q_default)
    dwell, root_state, transitions = get_history_statistics(T_aug, nstates, root=root)
    return (np.log(prior_root_distn[root_state]) - np.dot(dwell, total_rates) +
            special.xlogy(transitions, Q_default).sum())


def get_expm_augmented_tree(T, root, Q_default=None):
    """raoteh/sampler/_mjp_dense.py:328-359: every edge annotated with P = expm(Q t),
    all edges in one rt_expm_batched launch."""
    sched = TreeSchedule.from_nx(T, root)
    T_aug = nx.Graph()
    if sched.n == 1:
        return T_aug
    first = T[sched.nodes[sched.parent[1]]][sched.nodes[1]].get('Q', Q_default)
    check_square_dense(first)
    P, _ = _core.expm_edges(sched, T, first.shape[0], Q_default)
    for na, nb in nx.bfs_edges(T, root):
        T_aug.add_edge(na, nb, weight=T[na][nb]['weight'], P=P[sched.node_index[nb]])
    return T_aug


def get_likelihood(T, node_to_allowed_states, root, nstates, root_distn=None, Q_default=None):
    """raoteh/sampler/_mjp_dense.py:362-407 -> float likelihood (not log)."""
    if root not in T:
        raise ValueError('the specified root is not in the tree')
    T_aug = get_expm_augmented_tree(T, root, Q_default=Q_default)
    if len(T) == 1:
        T_aug.add_node(root)
    return _mcy_dense.get_likelihood(T_aug, root, nstates,
                                     node_to_allowed_states=node_to_allowed_states,
                                     root_distn=root_distn, P_default=None)


def _posterior(T, node_to_allowed_states, root, nstates, root_distn, Q_default):
    sched = TreeSchedule.from_nx(T, root)
    P, Qs = _core.expm_edges(sched, T, nstates, Q_default)
    ev = _core.Evaluation(sched, P, root_distn, nstates)
    mask = _core.mask_from_allowed(sched, node_to_allowed_states, nstates)
    mask = ev.support(mask, passes=3)
    ll, status, pmap = ev.upward_masks(mask)
    if status != 0:
        raise _util.NumericalZeroProb('the denominator is zero')
    D, J = ev.downward()
    M = ev.expectations(Qs, sched.length)
    return sched, Qs, D, J, M


def get_expected_history_statistics(T, node_to_allowed_states, root, nstates,
                                    root_distn=None, Q_default=None):
    """raoteh/sampler/_mjp_dense.py:410-539 -> (dict dwell, 1-D ndarray root posterior,
    nx.DiGraph of expected transition counts).

    Return types follow the code, not the docstring, of the reference (:536-539).
    Like the reference, the DiGraph has an entry for every (c, d) with
    Q[c, d] != 0 -- including the diagonal, whose weight Q[c,c] * E[dwell_c] is
    what the reference's loop at :513-533 produces.
    """
    if root not in T:
        raise ValueError('the specified root is not in the tree')
    sched, Qs, D, J, M = _posterior(T, node_to_allowed_states, root, nstates, root_distn, Q_default)
    dwell = defaultdict(float)
    trans = nx.DiGraph()
    for i in range(1, sched.n):
        Q = Qs[i]
        for sc in range(nstates):
            dwell[sc] += M[i, sc, sc]
        for sc in range(nstates):
            for sd in range(nstates):
                if not Q[sc, sd]:
                    continue
                if not trans.has_edge(sc, sd):
                    trans.add_edge(sc, sd, weight=0.0)
                trans[sc][sd]['weight'] += Q[sc, sd] * M[i, sc, sd]
    return dict(dwell), D[0], trans


def get_expected_ntransitions(T, node_to_allowed_states, root, nstates,
                              root_distn=None, Q_default=None, E=None):
    """Per-branch expected number of transitions weighted by E
    (examples/code2x3/extras.py:19-132) -> dict (na, nb) -> float."""
    if root not in T:
        raise ValueError('the specified root is not in the tree')
    if E is None:
        E = np.ones((nstates, nstates), dtype=float)
        np.fill_diagonal(E, 0)
    sched, Qs, D, J, M = _posterior(T, node_to_allowed_states, root, nstates, root_distn, Q_default)
    out = {}
    for na, nb in nx.bfs_edges(T, root):
        i = sched.node_index[nb]
        out[na, nb] = float((E * Qs[i] * M[i]).sum())
    return out


def differential_entropy_helper(Q, prior_root_distn, post_root_distn, post_dwell_times,
                                post_transitions):
    """raoteh/sampler/_mjp_dense.py:244-294"""
    check_square_dense(Q)
    check_square_dense(post_transitions)
    total_rates = get_total_rates(Q)
    diff_ent_init = -special.xlogy(post_root_distn, prior_root_distn).sum()
    diff_ent_dwell = post_dwell_times.dot(total_rates)
    diff_ent_trans = -special.xlogy(post_transitions, Q).sum()
    return diff_ent_init, diff_ent_dwell, diff_ent_trans
