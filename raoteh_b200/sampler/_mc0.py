"""
Sparse Markov-chain core with the reference's signatures (raoteh/sampler/_mc0.py):
distributions are dicts, transition matrices weighted nx.DiGraphs.  Adapters over
the dense CUDA path (see _mc0_dense).
"""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np

from . import _core, _sparse
from ._util import StructuralZeroProb, NumericalZeroProb, get_normalized_dict_distn

__all__ = []


def get_likelihood(root_pmap, root_distn=None):
    """raoteh/sampler/_mc0.py:202-252"""
    if (root_distn is not None) and not root_distn:
        raise StructuralZeroProb('no root state has nonzero prior likelihood')
    if root_pmap is None:
        raise ValueError('root_pmap is None')
    if not root_pmap:
        raise StructuralZeroProb('all root states give a subtree likelihood of zero')
    feasible = set(root_pmap)
    if root_distn is not None:
        feasible.intersection_update(set(root_distn))
    if not feasible:
        raise StructuralZeroProb('all root states have either zero prior likelihood '
                                 'or give a subtree likelihood of zero')
    if root_distn is not None:
        return sum(root_pmap[s] * root_distn[s] for s in feasible)
    return sum(root_pmap.values())


def _lower(T, root, node_to_pmap, root_distn, P_default):
    graphs = [T[a][b].get('P', P_default) for a, b in T.edges()]
    states = _sparse.state_space(graphs + [dict((s, 0) for p in node_to_pmap.values() for s in p)])
    index = dict((s, i) for i, s in enumerate(states))
    S = len(states)
    Td = nx.Graph()
    Td.add_nodes_from(T)
    for a, b in T.edges():
        P = T[a][b].get('P', P_default)
        if P is None:
            raise ValueError('no transition matrix is available')
        Td.add_edge(a, b, P=_sparse.dense_matrix(P, states, index))
    sched, Pd = _core.sched_and_P(Td, root, S, None)
    pmap = np.zeros((sched.n, S))
    for i, v in enumerate(sched.nodes):
        for s, x in node_to_pmap[v].items():
            pmap[i, index[s]] = x
    prior = None
    if root_distn is not None:
        prior = np.array([root_distn.get(s, 0.0) for s in states], dtype=float)
    return states, sched, Pd, pmap, prior


def get_node_to_distn(T, root, node_to_pmap, root_distn=None, P_default=None):
    """raoteh/sampler/_mc0.py:382-462 -> dict node -> dict state -> probability."""
    if len(T) == 1:
        return {root: get_normalized_dict_distn(node_to_pmap[root], root_distn)}
    get_normalized_dict_distn(node_to_pmap[root], root_distn)     # same exceptions as the reference
    states, sched, Pd, pmap, prior = _lower(T, root, node_to_pmap, root_distn, P_default)
    ev = _core.Evaluation(sched, Pd, prior, len(states))
    D, J = ev.downward_given_pmap(pmap)
    return dict((v, _sparse.vec_to_dict(D[i], states)) for i, v in enumerate(sched.nodes))


def get_joint_endpoint_distn(T, root, node_to_pmap, node_to_distn):
    """raoteh/sampler/_mc0.py:255-308 -> nx.Graph with sparse DiGraph attribute 'J'."""
    states, sched, Pd, pmap, _ = _lower(T, root, node_to_pmap, None, None)
    droot = np.array([node_to_distn[root].get(s, 0.0) for s in states])
    with np.errstate(divide='ignore', invalid='ignore'):
        prior = np.where(pmap[0] > 0, droot / pmap[0], 0.0)
    ev = _core.Evaluation(sched, Pd, prior, len(states))
    D, J = ev.downward_given_pmap(pmap)
    T_aug = nx.Graph()
    for i in range(1, sched.n):
        T_aug.add_edge(sched.nodes[sched.parent[i]], sched.nodes[i],
                       J=_sparse.sparse_matrix(J[i], states))
    return T_aug
