"""Shared test helpers: golden-fixture decoding and oracle/engine input builders."""
import json
import os

import networkx as nx
import numpy as np

from oracle import np_oracle
from raoteh_b200.lowering import TreeSchedule, allowed_sets_to_mask

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def case_tree(case):
    T = nx.Graph()
    for v in case['tree']['nodes']:
        T.add_node(v)
    for a, b, w in case['tree']['edges']:
        T.add_edge(a, b, weight=w)
    return T, case['tree']['root']


def case_allowed(case):
    return dict((int(k), set(v)) for k, v in case['allowed'].items())


def case_sched_mask(case):
    T, root = case_tree(case)
    sched = TreeSchedule.from_nx(T, root)
    S = case['nstates']
    mask = allowed_sets_to_mask(sched, case_allowed(case), S)
    return T, root, sched, mask


def oracle_obs_from_mask(mask, S):
    """mask uint64 [n] or [n,N] -> np_oracle.Obs"""
    m = np.asarray(mask, dtype=np.uint64)
    if m.ndim == 1:
        m = m[:, None]
    return np_oracle.Obs('mask', S, m.shape[1], mask=m)
