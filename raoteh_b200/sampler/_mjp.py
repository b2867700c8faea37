"""
Markov jump process likelihood and expectations with sparse (networkx) rate
matrices -- the reference's `raoteh.sampler._mjp` signatures
(raoteh/sampler/_mjp.py).  Sparse inputs are densified over the sorted state
set (as the reference itself does before scipy/pyfelscore, _mjp.py:468-478,
_linalg.py:72-90) and evaluated by the CUDA hot path.
"""
from __future__ import division, print_function, absolute_import

from collections import defaultdict

import networkx as nx
import numpy as np
from scipy import special

from . import _core, _mc0, _mcy, _sparse, _util
from ..lowering import TreeSchedule

__all__ = []


def get_total_rates(Q):
    """raoteh/sampler/_mjp.py:26-44 -> dict state -> total rate away."""
    total_rates = defaultdict(float)
    for a, b in Q.edges():
        total_rates[a] += Q[a][b]['weight']
    return dict(total_rates)


def get_conditional_transition_matrix(Q, total_rates=None):
    """raoteh/sampler/_mjp.py:47-71"""
    if total_rates is None:
        total_rates = get_total_rates(Q)
    P = nx.DiGraph()
    for a, b in Q.edges():
        P.add_edge(a, b, weight=Q[a][b]['weight'] / total_rates[a])
    return P


def get_history_dwell_times(T):
    """raoteh/sampler/_mjp.py:74-94"""
    dwell = defaultdict(float)
    for a, b in T.edges():
        dwell[T[a][b]['state']] += T[a][b]['weight']
    return dict(dwell)


def get_history_root_state_and_transitions(T, root=None):
    """raoteh/sampler/_mjp.py:97-147"""
    degrees = dict(T.degree())
    if root is None:
        root = _util.get_arbitrary_tip(T, degrees)
    root_states = [T[root][b]['state'] for b in T[root]]
    if len(set(root_states)) != 1:
        raise ValueError('the root does not have a well defined state')
    transitions = nx.DiGraph()
    successors = nx.dfs_successors(T, root)
    for a, b in nx.bfs_edges(T, root):
        if degrees[b] == 2:
            c = _util.get_first_element(successors[b])
            sa, sb = T[a][b]['state'], T[b][c]['state']
            if sa != sb:
                if transitions.has_edge(sa, sb):
                    transitions[sa][sb]['weight'] += 1
                else:
                    transitions.add_edge(sa, sb, weight=1)
    return root_states[0], transitions


def get_history_statistics(T, root=None):
    """raoteh/sampler/_mjp.py:150-183"""
    dwell = get_history_dwell_times(T)
    root_state, transitions = get_history_root_state_and_transitions(T, root=root)
    return dwell, root_state, transitions


def get_trajectory_log_likelihood(T_aug, root, prior_root_distn, Q_default):
    """raoteh/sampler/_mjp.py:186-252"""
    total_rates = get_total_rates(Q_default)
    dwell, root_state, transitions = get_history_statistics(T_aug, root=root)
    init_ll = np.log(prior_root_distn[root_state])
    dwell_ll = -sum(dwell[s] * total_rates.get(s, 0.0) for s in dwell)
    trans_ll = 0.0
    for sa, sb in transitions.edges():
        trans_ll += special.xlogy(transitions[sa][sb]['weight'], Q_default[sa][sb]['weight'])
    return init_ll + dwell_ll + trans_ll


def _lower(T, root, Q_default):
    """Dense per-edge rate matrices over the sorted union of states (_mjp.py:468-478)."""
    sched = TreeSchedule.from_nx(T, root)
    graphs = [Q_default]
    for i in range(1, sched.n):
        graphs.append(T[sched.nodes[sched.parent[i]]][sched.nodes[i]].get('Q', None))
    states = _sparse.state_space(graphs)
    index = dict((s, i) for i, s in enumerate(states))
    S = len(states)
    Td = nx.Graph()
    Td.add_nodes_from(T)
    reach = np.zeros((sched.n, S, S), dtype=bool)
    for i in range(1, sched.n):
        a, b = sched.nodes[sched.parent[i]], sched.nodes[i]
        Q = T[a][b].get('Q', Q_default)
        if Q is None:
            raise ValueError('no rate matrix is available for this edge')
        Qd = _sparse.dense_matrix(Q, states, index)
        np.fill_diagonal(Qd, 0.0)
        Qd -= np.diag(Qd.sum(axis=1))
        Td.add_edge(a, b, weight=T[a][b]['weight'], Q=Qd)
        reach[i] = _sparse.reachability(Q, states)
    return sched, states, index, Td, reach


def get_expm_augmented_tree(T, root, Q_default=None):
    """raoteh/sampler/_mjp.py:349-381: P = sparse_expm(Q, t) on every edge, as weighted
    DiGraphs holding the entries reachable in the digraph of Q (_linalg.py:83-89)."""
    T_aug = nx.Graph()
    if len(T) == 1:
        return T_aug
    sched, states, index, Td, reach = _lower(T, root, Q_default)
    P, _ = _core.expm_edges(sched, Td, len(states), None)
    for na, nb in nx.bfs_edges(T, root):
        i = sched.node_index[nb]
        T_aug.add_edge(na, nb, weight=T[na][nb]['weight'],
                       P=_sparse.sparse_matrix(P[i], states, reach[i]))
    return T_aug


def get_likelihood(T, node_to_allowed_states, root, root_distn=None, Q_default=None):
    """raoteh/sampler/_mjp.py:384-428 -> float likelihood."""
    if root not in T:
        raise ValueError('the specified root is not in the tree')
    T_aug = get_expm_augmented_tree(T, root, Q_default=Q_default)
    if len(T) == 1:
        T_aug.add_node(root)
    return _mcy.get_likelihood(T_aug, root, node_to_allowed_states=node_to_allowed_states,
                               root_distn=root_distn, P_default=None)


def get_expected_history_statistics(T, node_to_allowed_states, root, root_distn=None,
                                    Q_default=None):
    """raoteh/sampler/_mjp.py:431-594 -> (dict state -> expected dwell, dict state ->
    root posterior, nx.DiGraph of expected transition counts)."""
    if root not in T:
        raise ValueError('the specified root is not in the tree')
    sched, states, index, Td, reach = _lower(T, root, Q_default)
    S = len(states)
    P, Qs = _core.expm_edges(sched, Td, S, None)
    P = P * reach                     # sparse_expm keeps only reachable entries
    prior = None
    if root_distn is not None:
        prior = np.array([root_distn.get(s, 0.0) for s in states], dtype=float)
    ev = _core.Evaluation(sched, P, prior, S)
    mask = _core.mask_from_allowed(sched, _sparse.allowed_to_index(node_to_allowed_states, index), S)
    mask = ev.support(mask, passes=3)
    ll, status, pmap = ev.upward_masks(mask)
    if status != 0:
        # same exception the reference raises from _mc0.get_node_to_distn
        _util.get_normalized_dict_distn(_sparse.vec_to_dict(pmap[0], states), root_distn)
        raise _util.NumericalZeroProb('the denominator is zero')
    D, J = ev.downward()
    M = ev.expectations(Qs, sched.length)
    dwell = defaultdict(float)
    trans = nx.DiGraph()
    for i in range(1, sched.n):
        a, b = sched.nodes[sched.parent[i]], sched.nodes[i]
        Q = T[a][b].get('Q', Q_default)
        for sc_index, sc in enumerate(states):
            dwell[sc] += M[i, sc_index, sc_index]
        for sc, sd in Q.edges():
            if not trans.has_edge(sc, sd):
                trans.add_edge(sc, sd, weight=0.0)
            trans[sc][sd]['weight'] += Q[sc][sd]['weight'] * M[i, index[sc], index[sd]]
    return dict(dwell), _sparse.vec_to_dict(D[0], states), trans


def differential_entropy_helper(Q, prior_root_distn, post_root_distn, post_dwell_times,
                                post_transitions):
    """raoteh/sampler/_mjp.py:255-306: the three contributions (initial state, dwell times,
    transitions) to the expected negative log-likelihood of a trajectory, from posterior
    expectations such as those of get_expected_history_statistics."""
    from scipy import special
    total_rates = get_total_rates(Q)
    diff_ent_init = 0.0
    for state, prob in post_root_distn.items():
        diff_ent_init -= special.xlogy(prob, prior_root_distn[state])
    diff_ent_dwell = 0.0
    for s in set(total_rates) & set(post_dwell_times):
        diff_ent_dwell += post_dwell_times[s] * total_rates[s]
    diff_ent_trans = 0.0
    for sa in set(Q) & set(post_transitions):
        for sb in set(Q[sa]) & set(post_transitions[sa]):
            diff_ent_trans -= special.xlogy(post_transitions[sa][sb]['weight'], Q[sa][sb]['weight'])
    return diff_ent_init, diff_ent_dwell, diff_ent_trans
