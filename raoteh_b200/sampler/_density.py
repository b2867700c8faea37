"""
networkx <-> ndarray lowering helpers with the reference's names and conventions
(raoteh/sampler/_density.py): dense rate matrices carry their diagonal, trees go
to boolean CSR over nodes in a requested (DFS-preorder) order, edge-specific
matrices are indexed by the child node.
"""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np

from ..lowering import check_square_dense

__all__ = ['check_square_dense', 'digraph_to_bool_csr', 'get_esd_transitions',
           'dict_to_numpy_array', 'rate_matrix_to_numpy_array']


def rate_matrix_to_numpy_array(Q_sparse, **kwargs):
    """raoteh/sampler/_density.py:32-54: rows sum to zero."""
    pre = np.asarray(nx.to_numpy_array(Q_sparse, **kwargs), dtype=float)
    return pre - np.diag(pre.sum(axis=1))


def dict_to_numpy_array(d, nodelist=None):
    """raoteh/sampler/_density.py:57-75"""
    if nodelist is None:
        nodelist = tuple(d)
    return np.array([d.get(n, 0) for n in nodelist], dtype=float)


def digraph_to_bool_csr(G, ordered_nodes):
    """raoteh/sampler/_density.py:104-140"""
    node_to_index = dict((n, i) for i, n in enumerate(ordered_nodes))
    indices, indptr = [], [0]
    for na in ordered_nodes:
        if na in G:
            for nb in G[na]:
                indices.append(node_to_index[nb])
        indptr.append(len(indices))
    return np.array(indices, dtype=int), np.array(indptr, dtype=int)


def get_esd_transitions(G, preorder_nodes, nstates, P_default=None):
    """raoteh/sampler/_density.py:143-180: (nnodes, S, S) indexed by the child."""
    nnodes = len(preorder_nodes)
    if nnodes != G.number_of_nodes():
        raise ValueError('the number of nodes is inconsistent')
    node_to_index = dict((n, i) for i, n in enumerate(preorder_nodes))
    out = np.zeros((nnodes, nstates, nstates), dtype=float)
    for na in preorder_nodes:
        if na in G:
            for nb in G[na]:
                P = G[na][nb].get('P', P_default)
                check_square_dense(P)
                out[node_to_index[nb]] = P
    return out
