"""
Observation type y (node -> set of allowed states) with dense transition
matrices, the reference's signatures (raoteh/sampler/_mcy_dense.py).  The three
pyfelscore calls of `_esd_get_node_to_pmap` (:233-299) -- backward support,
forward support, pruning -- run as rt_support_sets + rt_prune_loglik.
"""
from __future__ import division, print_function, absolute_import

import numpy as np

from . import _core, _mc0_dense, _util

__all__ = []


def _single_node_pmap(T, root, nstates, node_to_allowed_states):
    _util._check_root(T, root)
    allowed = set(range(nstates))
    if node_to_allowed_states is not None:
        allowed &= set(node_to_allowed_states[root])
    return np.array([1 if s in allowed else 0 for s in range(nstates)], dtype=float)


def _evaluate(T, root, nstates, node_to_allowed_states, root_distn, P_default):
    sched, P = _core.sched_and_P(T, root, nstates, P_default)
    ev = _core.Evaluation(sched, P, root_distn, nstates)
    mask = _core.mask_from_allowed(sched, node_to_allowed_states, nstates)
    mask = ev.support(mask, passes=3)
    ll, status, pmap = ev.upward_masks(mask)
    return sched, ev, mask, pmap, status


def get_node_to_pmap(T, root, nstates, node_to_allowed_states=None, P_default=None,
                     node_to_set=None):
    """raoteh/sampler/_mcy_dense.py:302-354 -> dict node -> 1d ndarray of subtree likelihoods."""
    if len(T) == 1 and P_default is not None:
        return {root: _single_node_pmap(T, root, nstates, node_to_allowed_states)}
    best = node_to_set if node_to_set is not None else node_to_allowed_states
    sched, ev, mask, pmap, status = _evaluate(T, root, nstates, best, None, P_default)
    return dict((v, pmap[i]) for i, v in enumerate(sched.nodes))


def kitchen_sink(T, root, nstates, node_to_allowed_states=None, root_distn=None, P_default=None):
    """raoteh/sampler/_mcy_dense.py:57-127 -> (node_to_pmap, node_to_distn, edge_to_joint_distn)."""
    _util._check_root(T, root)
    if len(T) == 1:
        root_pmap = _single_node_pmap(T, root, nstates, node_to_allowed_states)
        w = root_pmap if root_distn is None else root_pmap * root_distn
        return {root: root_pmap}, {root: w / w.sum()}, {}
    sched, ev, mask, pmap, status = _evaluate(T, root, nstates, node_to_allowed_states,
                                              root_distn, P_default)
    if status != 0:
        raise _util.NumericalZeroProb('the denominator is zero')
    D, J = ev.downward()
    node_to_pmap = dict((v, pmap[i]) for i, v in enumerate(sched.nodes))
    node_to_distn = dict((v, D[i]) for i, v in enumerate(sched.nodes))
    edge_to_joint = dict(((sched.nodes[sched.parent[i]], sched.nodes[i]), J[i])
                         for i in range(1, sched.n))
    return node_to_pmap, node_to_distn, edge_to_joint


def get_likelihood(T, root, nstates, node_to_allowed_states=None, root_distn=None, P_default=None):
    """raoteh/sampler/_mcy_dense.py:433-493; raises StructuralZeroProb on empty support."""
    if len(T) == 1:
        _util._check_root(T, root)
        allowed = node_to_allowed_states[root]
        if not allowed:
            raise _util.StructuralZeroProb('the tree has only a single node, '
                                           'and no state is allowed for the root')
        if root_distn is None:
            return 1
        pos = set(s for s in allowed if root_distn[s])
        if not pos:
            raise _util.StructuralZeroProb(
                'the tree has only a single node, and every state with positive prior '
                'probability at the root is disallowed by a node state constraint')
        return sum(root_distn[s] for s in pos)
    node_to_pmap = get_node_to_pmap(T, root, nstates,
                                    node_to_allowed_states=node_to_allowed_states,
                                    P_default=P_default)
    return _mc0_dense.get_likelihood(node_to_pmap[root], root_distn=root_distn)
