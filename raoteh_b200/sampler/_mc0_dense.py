"""
Dense Markov-chain core on trees with the reference's signatures
(raoteh/sampler/_mc0_dense.py): root combine, downward pass (node marginals)
and per-edge joint endpoint distributions, computed by the CUDA kernels
(rt_posterior_stats, rt_joint_distn) on a batch of one site.
"""
from __future__ import division, print_function, absolute_import

import warnings

import networkx as nx
import numpy as np

from . import _core
from ._util import StructuralZeroProb, NumericalZeroProb, get_normalized_ndarray_distn

__all__ = []


def get_likelihood(root_pmap, root_distn=None):
    """raoteh/sampler/_mc0_dense.py:147-212 (root combine; same exceptions)."""
    if root_distn is not None:
        if root_pmap.shape != root_distn.shape:
            raise ValueError('root shape mismatch: %s %s' % (root_pmap.shape, root_distn.shape))
        prior_feasible = set(s for s, p in enumerate(root_distn) if p)
        if not prior_feasible:
            raise StructuralZeroProb('no root state has nonzero prior likelihood')
    if root_pmap is None:
        raise ValueError('root_pmap is None')
    if root_pmap.min() < 0:
        warnings.warn('root_pmap should have non-negative entries '
                      'but found minimum entry %s' % root_pmap.min())
        root_pmap = np.maximum(root_pmap, 0)
    if not root_pmap.sum():
        raise StructuralZeroProb('all root states give a subtree likelihood of zero')
    feasible = set(s for s, p in enumerate(root_pmap) if p)
    if root_distn is not None:
        feasible.intersection_update(prior_feasible)
    if not feasible:
        raise StructuralZeroProb('all root states have either zero prior likelihood '
                                 'or give a subtree likelihood of zero')
    if root_distn is not None:
        return root_distn.dot(root_pmap)
    return root_pmap.sum()


def _pmap_array(sched, node_to_pmap, nstates):
    pmap = np.zeros((sched.n, nstates))
    for i, v in enumerate(sched.nodes):
        p = np.asarray(node_to_pmap[v], dtype=float)
        if p.shape[0] != nstates:
            raise ValueError('inconsistent pmap')
        pmap[i] = p
    return pmap


def get_node_to_distn(T, root, node_to_pmap, nstates, root_distn=None, P_default=None):
    """raoteh/sampler/_mc0_dense.py:400-489 -> dict node -> 1d ndarray."""
    if P_default is not None:
        _core.check_square_dense(P_default)
    if root_distn is not None and root_distn.shape[0] != nstates:
        raise ValueError('inconsistent root distribution')
    if len(T) == 1:
        return {root: get_normalized_ndarray_distn(np.asarray(node_to_pmap[root], float), root_distn)}
    sched, P = _core.sched_and_P(T, root, nstates, P_default)
    ev = _core.Evaluation(sched, P, root_distn, nstates)
    D, J = ev.downward_given_pmap(_pmap_array(sched, node_to_pmap, nstates))
    return dict((v, D[i]) for i, v in enumerate(sched.nodes))


get_node_to_distn_esd = get_node_to_distn   # raoteh/sampler/_mc0_dense.py:344 (pyfelscore variant)


def get_joint_endpoint_distn(T, root, node_to_pmap, node_to_distn, nstates):
    """raoteh/sampler/_mc0_dense.py:217-270 -> nx.Graph with edge attribute 'J'."""
    sched, P = _core.sched_and_P(T, root, nstates, None)
    distn = node_to_distn[root]
    if distn.shape[0] != nstates:
        raise Exception('nstates inconsistency')
    pmap = _pmap_array(sched, node_to_pmap, nstates)
    # the reference takes node_to_distn as given; the root row fixes the rest
    pmap_root_w = pmap[0].copy()
    with np.errstate(divide='ignore', invalid='ignore'):
        prior = np.where(pmap_root_w > 0, np.asarray(distn, float) / pmap_root_w, 0.0)
    ev = _core.Evaluation(sched, P, prior, nstates)
    D, J = ev.downward_given_pmap(pmap)
    T_aug = nx.Graph()
    for i in range(1, sched.n):
        T_aug.add_edge(sched.nodes[sched.parent[i]], sched.nodes[i], J=J[i])
    return T_aug
