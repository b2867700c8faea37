"""
Observation type y with sparse transition matrices -- the reference's
`raoteh.sampler._mcy` signatures (raoteh/sampler/_mcy.py).  The sparse inputs
are densified exactly as the reference does before its own native calls
(_mcy.py:473-560: CSR tree + 3-D P array + state mask) and run on the GPU.
"""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np

from . import _core, _mc0, _sparse, _util

__all__ = []


def _lower(T, root, node_to_allowed_states, P_default):
    edges = list(nx.bfs_edges(T, root))
    any_custom = any('P' in T[a][b] for a, b in edges)
    all_custom = all('P' in T[a][b] for a, b in edges)
    if (P_default is None) and (not all_custom):
        raise ValueError('expected a custom transition on each edge '
                         'when a default transition matrix is not available')
    graphs = [T[a][b].get('P', P_default) for a, b in edges]
    states = _sparse.state_space(graphs)
    index = dict((s, i) for i, s in enumerate(states))
    S = len(states)
    Td = nx.Graph()
    Td.add_nodes_from(T)
    for (a, b), P in zip(edges, graphs):
        Td.add_edge(a, b, P=_sparse.dense_matrix(P, states, index))
    sched, Pd = _core.sched_and_P(Td, root, S, None)
    mask = _core.mask_from_allowed(sched, _sparse.allowed_to_index(node_to_allowed_states, index), S)
    return states, sched, Pd, mask


def _mask_to_sets(sched, mask, states):
    out = {}
    for i, v in enumerate(sched.nodes):
        out[v] = set(states[s] for s in range(len(states)) if (int(mask[i]) >> s) & 1)
    return out


def get_node_to_pset(T, root, node_to_allowed_states=None, P_default=None):
    """raoteh/sampler/_mcy.py:348-394: states with positive subtree likelihood
    (backward pass only)."""
    if len(T) == 1:
        return {root: set(node_to_allowed_states[root])}
    states, sched, Pd, mask = _lower(T, root, node_to_allowed_states, P_default)
    ev = _core.Evaluation(sched, Pd, None, len(states))
    return _mask_to_sets(sched, ev.support(mask, passes=1), states)


def get_node_to_set(T, root, node_to_allowed_states=None, P_default=None):
    """raoteh/sampler/_mcy.py:323-345: backward then forward support."""
    if len(T) == 1:
        return {root: set(node_to_allowed_states[root])}
    states, sched, Pd, mask = _lower(T, root, node_to_allowed_states, P_default)
    ev = _core.Evaluation(sched, Pd, None, len(states))
    return _mask_to_sets(sched, ev.support(mask, passes=3), states)


def get_node_to_pmap(T, root, node_to_allowed_states=None, P_default=None, node_to_set=None):
    """raoteh/sampler/_mcy.py:563-608 -> dict node -> dict state -> subtree likelihood
    (only states in the node's support set appear)."""
    if len(T) == 1 and root in T:
        allowed = node_to_allowed_states[root] if node_to_set is None else node_to_set[root]
        return {root: dict((s, 1.0) for s in allowed)}
    best = node_to_set if node_to_set is not None else node_to_allowed_states
    states, sched, Pd, mask = _lower(T, root, best, P_default)
    ev = _core.Evaluation(sched, Pd, None, len(states))
    mask = ev.support(mask, passes=3)
    ll, status, pmap = ev.upward_masks(mask)
    out = {}
    for i, v in enumerate(sched.nodes):
        support = [(int(mask[i]) >> s) & 1 for s in range(len(states))]
        out[v] = _sparse.vec_to_dict(pmap[i], states, support)
    return out


def get_likelihood(T, root, node_to_allowed_states=None, root_distn=None, P_default=None):
    """raoteh/sampler/_mcy.py:685-746"""
    if len(T) == 1:
        _util._check_root(T, root)
        allowed = node_to_allowed_states[root]
        if not allowed:
            raise _util.StructuralZeroProb('the tree has only a single node, '
                                           'and no state is allowed for the root')
        if root_distn is None:
            return 1
        pos = set(allowed) & set(s for s, p in root_distn.items() if p)
        if not pos:
            raise _util.StructuralZeroProb(
                'the tree has only a single node, and every state with positive prior '
                'probability at the root is disallowed by a node state constraint')
        return sum(root_distn[s] for s in pos)
    node_to_pmap = get_node_to_pmap(T, root, node_to_allowed_states=node_to_allowed_states,
                                    P_default=P_default)
    return _mc0.get_likelihood(node_to_pmap[root], root_distn=root_distn)
