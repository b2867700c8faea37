"""Uniformization helpers with the reference's signatures
(raoteh/sampler/_sample_mjp.py:72-117, dense twin _sample_mjp_dense.py:72-112)."""
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
