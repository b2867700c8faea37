"""
Hard-observation wrappers of the chain samplers (raoteh/sampler/_sample_mcx.py): every observed
node has exactly one allowed state.
"""
from __future__ import division, print_function, absolute_import

from . import _sample_mcy, _sampler, _sparse  # noqa: F401

__all__ = []


def _allowed(node_to_state):
    return dict((v, {s}) for v, s in (node_to_state or {}).items())


def resample_states(T, root, node_to_state=None, root_distn=None, P_default=None, seed=None):
    """raoteh/sampler/_sample_mcx.py:103-170"""
    return _sample_mcy.resample_states(T, root, node_to_allowed_states=_allowed(node_to_state),
                                       root_distn=root_distn, P_default=P_default, seed=seed)


def resample_edge_states(T, root, event_nodes, node_to_state=None, root_distn=None,
                         P_default=None, seed=None):
    """raoteh/sampler/_sample_mcx.py:173-261"""
    return _sample_mcy.resample_edge_states(T, root, P_default, event_nodes,
                                            node_to_allowed_states=_allowed(node_to_state),
                                            root_distn=root_distn, seed=seed)


def get_feasible_history(T, node_to_state, root=None, root_distn=None, P_default=None):
    """raoteh/sampler/_sample_mcx.py:20-100: an arbitrary feasible history under the transition
    matrix P_default (0, 1, 3, 7, ... equally spaced events per edge until FFBS succeeds)."""
    from ._util import get_first_element
    if root is None:
        root = get_first_element(node_to_state) if node_to_state else get_first_element(T)
    allowed = dict((v, set(P_default)) for v in T)
    allowed.update(_allowed(node_to_state))
    return _sampler.get_restricted_feasible_history(T, P_default, allowed, root, root_distn=root_distn)
