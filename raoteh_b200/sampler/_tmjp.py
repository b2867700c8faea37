"""
Sparse tolerance-model helpers (raoteh/sampler/_tmjp.py): the model container and the
functions around the hot path whose arguments are nx.DiGraph rate matrices and dict
distributions.  Numerical work is delegated to the dense mirror (_tmjp_dense) and so to
the CUDA library.
"""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np

from . import _sparse, _tmjp_dense

__all__ = []


def get_tolerance_rate_matrix(rate_off, rate_on):
    """raoteh/sampler/_tmjp.py:352-373"""
    Q = nx.DiGraph()
    if rate_on:
        Q.add_edge(0, 1, weight=rate_on)
    if rate_off:
        Q.add_edge(1, 0, weight=rate_off)
    return Q


def get_tolerance_distn(rate_off, rate_on):
    """raoteh/sampler/_tmjp.py:376-403"""
    if (rate_off < 0) or (rate_on < 0):
        raise ValueError('rates must be non-negative')
    total = rate_off + rate_on
    if total <= 0:
        raise ValueError('the total tolerance rate must be positive')
    distn = {}
    if rate_off:
        distn[0] = rate_off / total
    if rate_on:
        distn[1] = rate_on / total
    return distn


class CompoundToleranceModel(object):
    """raoteh/sampler/_tmjp.py:30-64 (sparse inputs; compound attributes on demand)."""

    def __init__(self, Q_primary, primary_distn, primary_to_part, rate_on, rate_off):
        self.Q_primary = Q_primary
        self.primary_distn = primary_distn
        self.primary_to_part = primary_to_part
        self.rate_on = rate_on
        self.rate_off = rate_off
        self.nprimary = len(primary_to_part)
        self.nparts = len(set(primary_to_part.values()))
        self.ncompound = int(np.ldexp(self.nprimary, self.nparts))
        self.tolerance_distn = get_tolerance_distn(rate_off, rate_on)
        self.Q_compound = None
        self.compound_distn = None
        self.compound_to_primary = None
        self.compound_to_tolerances = None

    def _dense(self):
        states = sorted(self.primary_to_part)
        index = dict((s, i) for i, s in enumerate(states))
        Q = _sparse.dense_matrix(self.Q_primary, states, index)
        Q -= np.diag(Q.sum(axis=1))
        distn = np.array([self.primary_distn.get(s, 0.0) for s in states], dtype=float)
        part = dict((index[s], c) for s, c in self.primary_to_part.items())
        return states, index, Q, distn, part


def get_primary_proposal_rate_matrix(Q_primary, primary_to_part, tolerance_distn):
    """raoteh/sampler/_tmjp.py:905-958"""
    Q_proposal = nx.DiGraph()
    for sa, sb in Q_primary.edges():
        rate = Q_primary[sa][sb]['weight']
        if primary_to_part[sa] == primary_to_part[sb]:
            Q_proposal.add_edge(sa, sb, weight=rate)
        elif 1 in tolerance_distn:
            Q_proposal.add_edge(sa, sb, weight=rate * tolerance_distn[1])
    return Q_proposal


def get_tolerance_summary(ctm, T_primary, root, disease_data=None):
    """raoteh/sampler/_tmjp.py:613-741 -> the seven tolerance expectations."""
    states, index, Q, distn, part = ctm._dense()
    T_dense = nx.Graph()
    for a, b, d in T_primary.edges(data=True):
        T_dense.add_edge(a, b, weight=d['weight'], state=index[d['state']])
    return _tmjp_dense.get_tolerance_summary(part, ctm.rate_on, ctm.rate_off, Q, T_dense, root,
                                             disease_data=disease_data)


def get_tolerance_ll_contribs(rate_on, rate_off, total_tree_length, *summary):
    """raoteh/sampler/_tmjp.py:744-812"""
    return _tmjp_dense.get_tolerance_ll_contribs(rate_on, rate_off, total_tree_length, *summary)


def get_tolerance_process_log_likelihood(ctm, T_primary, root):
    """raoteh/sampler/_tmjp.py:406-490: compound log-likelihood of a primary trajectory with the
    tolerance histories integrated out (sparse inputs; evaluated by the dense twin)."""
    if root is None:
        raise ValueError('unspecified root')
    if root not in T_primary:
        raise ValueError('the specified root is not a node in the tree')
    states, index, Q, distn, part = ctm._dense()
    T_dense = nx.Graph()
    for a, b, d in T_primary.edges(data=True):
        T_dense.add_edge(a, b, weight=d['weight'], state=index[d['state']])
    if len(T_primary) == 1:
        T_dense.add_node(root)
    return _tmjp_dense.get_tolerance_process_log_likelihood(
        Q, part, T_dense, ctm.rate_off, ctm.rate_on, distn, root)


def differential_entropy_helper(ctm, post_root_distn, post_dwell_times, post_transitions):
    """raoteh/sampler/_tmjp.py:217-349: sparse inputs (dicts over compound states, DiGraph of
    expected transition counts) -> CompoundNegLL, through the dense twin on a dense compound model
    with the same state numbering (itertools.product order, raoteh/sampler/_tmjp.py:75-83)."""
    states, index, Q, distn, part = ctm._dense()
    dense = _tmjp_dense.CompoundToleranceModel(Q, distn, part, ctm.rate_on, ctm.rate_off)
    dense.init_compound()
    # the sparse model has transitions only between states of positive prior probability
    # (raoteh/sampler/_tmjp.py:117-119); the dense one also between the formal infeasible states
    ok = np.asarray(dense.compound_distn) > 0
    Qc = np.where(ok[:, None] & ok[None, :], dense.Q_compound, 0.0)
    np.fill_diagonal(Qc, 0.0)
    dense.Q_compound = Qc - np.diag(Qc.sum(axis=1))
    n = dense.ncompound
    root = np.zeros(n)
    for k, v in post_root_distn.items():
        root[k] = v
    dwell = np.zeros(n)
    for k, v in post_dwell_times.items():
        dwell[k] = v
    trans = np.zeros((n, n))
    for a, b, d in post_transitions.edges(data=True):
        if a != b and Qc[a, b] > 0:        # only transitions the sparse model has (:329-331)
            trans[a, b] = d['weight']
    return _tmjp_dense.differential_entropy_helper(dense, root, dwell, trans)
