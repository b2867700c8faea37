"""
Observation type z (node -> {state: likelihood}), the reference's
`raoteh.sampler._mcz` signatures (raoteh/sampler/_mcz.py).  The emission
likelihood multiplies the subtree likelihood at every node (:159-160); on the
GPU this is the OBS_DENSE form of rt_prune_loglik.
"""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np

from . import _core, _mc0, _mcy, _sparse

__all__ = []


def _allowed(node_to_state_to_likelihood):
    if node_to_state_to_likelihood is None:
        return None
    return dict((n, set(d)) for n, d in node_to_state_to_likelihood.items())


def get_node_to_set(T, root, node_to_state_to_likelihood=None, P_default=None):
    """raoteh/sampler/_mcz.py:29-41"""
    return _mcy.get_node_to_set(T, root, node_to_allowed_states=_allowed(node_to_state_to_likelihood),
                                P_default=P_default)


def get_node_to_pset(T, root, node_to_state_to_likelihood=None, P_default=None):
    """raoteh/sampler/_mcz.py:44-91"""
    return _mcy.get_node_to_pset(T, root, node_to_allowed_states=_allowed(node_to_state_to_likelihood),
                                 P_default=P_default)


def get_node_to_pmap(T, root, node_to_state_to_likelihood=None, P_default=None, node_to_set=None):
    """raoteh/sampler/_mcz.py:94-166; requires an entry for every node (:159)."""
    if node_to_set is None:
        node_to_set = get_node_to_set(T, root, node_to_state_to_likelihood, P_default)
    if len(T) == 1:
        return {root: dict((s, 1.0 * node_to_state_to_likelihood[root][s]) for s in node_to_set[root])}
    states, sched, Pd, mask = _mcy._lower(T, root, node_to_set, P_default)
    index = dict((s, i) for i, s in enumerate(states))
    S = len(states)
    lik = np.zeros((sched.n, S))
    for i, v in enumerate(sched.nodes):
        for s in node_to_set[v]:
            lik[i, index[s]] = node_to_state_to_likelihood[v][s]
    ev = _core.Evaluation(sched, Pd, None, S)
    ll, status, pmap = ev.upward_dense(lik)
    out = {}
    for i, v in enumerate(sched.nodes):
        support = [(int(mask[i]) >> s) & 1 for s in range(S)]
        out[v] = _sparse.vec_to_dict(pmap[i], states, support)
    return out


def get_likelihood(T, root, node_to_state_to_likelihood=None, root_distn=None, P_default=None):
    """raoteh/sampler/_mcz.py:169-211.  The reference's body refers to an undefined
    name (its parameter is called node_to_allowed_states, :170 vs :205) and cannot
    run; this is the evident intent."""
    node_to_pmap = get_node_to_pmap(T, root,
                                    node_to_state_to_likelihood=node_to_state_to_likelihood,
                                    P_default=P_default)
    return _mc0.get_likelihood(node_to_pmap[root], root_distn=root_distn)
