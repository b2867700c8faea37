"""Jukes-Cantor closed forms used as known answers (Tataru & Hobolth 2011), with
the reference's names (raoteh/sampler/_conditional_expectation.py:15-46)."""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np

__all__ = []


def get_jukes_cantor_rate_matrix(n=4):
    Q = nx.DiGraph()
    for i in range(n):
        for j in range(n):
            if i != j:
                Q.add_edge(i, j, weight=1.0 / (n - 1))
    return Q


def get_jukes_cantor_probability(i, j, t, n=4):
    p = np.exp(-(n * t) / (n - 1))
    return (1 + p * (n - 1)) / n if i == j else (1 - p) / n


def get_jukes_cantor_interaction(a, b, c, d, t, n=4):
    p = np.exp(-(n * t) / (n - 1))
    pm1 = np.expm1(-(n * t) / (n - 1))
    if a != c and d != b:
        x = t * p + pm1 * 2 * (n - 1) / n
    elif a == c and d == b:
        x = (n - 1) * (n - 1) * t * p - pm1 * 2 * (n - 1) * (n - 1) / n
    else:
        x = -(n - 1) * t * p - pm1 * (n - 2) * (n - 1) / n
    return (t + x) / (n * n)
