"""Sparse (networkx / dict) <-> dense adapters shared by the sparse mirror modules.
State labels are arbitrary sortable hashables (raoteh/sampler/_mcy.py:198-204)."""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np


def state_space(graphs):
    states = set()
    for G in graphs:
        if G is not None:
            states.update(G)
    return sorted(states)


def dense_matrix(G, states, index=None):
    """Weighted DiGraph -> dense [S,S] (zeros where no edge)."""
    if index is None:
        index = dict((s, i) for i, s in enumerate(states))
    M = np.zeros((len(states), len(states)), dtype=float)
    for a, b, d in G.edges(data=True):
        if a in index and b in index:
            M[index[a], index[b]] = d['weight']
    return M


def sparse_matrix(M, states, pattern=None):
    """dense -> weighted DiGraph keeping entries where `pattern` (bool) or M != 0."""
    G = nx.DiGraph()
    S = len(states)
    for i in range(S):
        for j in range(S):
            keep = pattern[i, j] if pattern is not None else (M[i, j] != 0)
            if keep:
                G.add_edge(states[i], states[j], weight=float(M[i, j]))
    return G


def reachability(Q, states):
    """bool[S,S]: j reachable from i in the digraph of Q (i reaches itself),
    the sparsity rule of raoteh/sampler/_linalg.py:83-89."""
    index = dict((s, i) for i, s in enumerate(states))
    S = len(states)
    R = np.zeros((S, S), dtype=bool)
    for s in states:
        if s in Q:
            for t in nx.descendants(Q, s) | {s}:
                if t in index:
                    R[index[s], index[t]] = True
    return R


def allowed_to_index(node_to_allowed_states, index):
    if node_to_allowed_states is None:
        return None
    out = {}
    for v, allowed in node_to_allowed_states.items():
        out[v] = set(index[s] for s in allowed if s in index)
    return out


def vec_to_dict(v, states, support=None):
    if support is None:
        return dict((states[i], float(x)) for i, x in enumerate(v) if x)
    return dict((states[i], float(v[i])) for i in range(len(states)) if support[i])
