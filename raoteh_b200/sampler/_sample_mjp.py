"""Uniformization helpers with the reference's signatures
(raoteh/sampler/_sample_mjp.py:19-117, dense twin _sample_mjp_dense.py:21-112).

`resample_poisson` is the reference's stand-alone Poisson step on an nx trajectory; inside the
batched sweeps the same step runs in the kernels (rt_raoteh.cu: hazard-space thinning at rate
omega - q_s per segment)."""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np

from . import _mjp

__all__ = []


def get_uniformized_transition_matrix(Q, uniformization_factor=None, omega=None):
    """Sparse B = I + Q/omega as a weighted DiGraph (raoteh/sampler/_sample_mjp.py:72-117)."""
    if (uniformization_factor is not None) and (omega is not None):
        raise ValueError('the uniformization factor and omega should not both be provided')
    total_rates = _mjp.get_total_rates(Q)
    if omega is None:
        if uniformization_factor is None:
            uniformization_factor = 2
        omega = uniformization_factor * max(total_rates.values())
    P = nx.DiGraph()
    for a in Q:
        if Q[a]:
            P.add_edge(a, a, weight=1.0 - total_rates[a] / omega)
            for b in Q[a]:
                P.add_edge(a, b, weight=Q[a][b]['weight'] / omega)
    return P


def get_uniformized_transition_matrix_dense(Q, uniformization_factor=None, omega=None):
    """Dense twin (raoteh/sampler/_sample_mjp_dense.py:72-112)."""
    if (uniformization_factor is not None) and (omega is not None):
        raise ValueError('the uniformization factor and omega should not both be provided')
    if omega is None:
        if uniformization_factor is None:
            uniformization_factor = 2
        omega = uniformization_factor * np.max(-np.diag(Q))
    return np.eye(Q.shape[0]) + Q / omega


def resample_poisson(T, state_to_rate, root=None):
    """raoteh/sampler/_sample_mjp.py:19-69 (dense twin _sample_mjp_dense.py:21-69, where
    `state_to_rate` is a 1-D array): drop a Poisson process of rate state_to_rate[state] onto every
    edge of the trajectory T (edges carry `weight` and `state`); returns the tree with the new
    degree-two nodes (ids above max(T)) and WITHOUT state annotation.  Uses numpy's global
    generator like the reference (sequential exponential gaps until the edge is used up)."""
    from . import _util
    if root is None:
        root = _util.get_first_element(T)
    next_node = max(T) + 1
    weighted_edges = []
    for a, b in nx.bfs_edges(T, root):
        weight = T[a][b]['weight']
        rate = state_to_rate[T[a][b]['state']]
        prev_node, total_dwell = a, 0.0
        while rate > 0:
            dwell = np.random.exponential(scale=1.0 / rate)
            if total_dwell + dwell > weight:
                break
            total_dwell += dwell
            weighted_edges.append((prev_node, next_node, dwell))
            prev_node = next_node
            next_node += 1
        weighted_edges.append((prev_node, b, weight - total_dwell))
    T_out = nx.Graph()
    T_out.add_weighted_edges_from(weighted_edges)
    return T_out


resample_poisson_dense = resample_poisson
