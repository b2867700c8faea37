"""
Samplers of discrete-time Markov chain states on trees (raoteh/sampler/_sample_mcy.py:19-187,
thin hard-observation wrappers in _sample_mcx.py): node states given a transition matrix per
edge step, and edge states on a tree whose `event_nodes` are the only places where the state
may change.  Both are forward-filter backward-sampling on a (chunk) tree; here they are the
Rao-Teh kernel run over a fixed set of candidate events with the chain's transition matrix in
place of the uniformized one and no virtual events.
"""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np

from . import _sparse
from ._util import StructuralZeroProb
from .. import engine, _native
from ..lowering import TreeSchedule
from ..raoteh import RaoTehChains

__all__ = []


def _states_and_matrix(P_default, root_distn, node_to_allowed_states):
    if P_default is None:
        raise NotImplementedError('per-edge transition matrices are not supported: pass P_default')
    states = set(P_default)
    if root_distn is not None:
        states |= set(root_distn)
    if node_to_allowed_states:
        for allowed in node_to_allowed_states.values():
            states |= set(allowed)
    states = sorted(states)
    if len(states) < 2:
        states = states + [object()] if states else [0, 1]
    index = dict((s, i) for i, s in enumerate(states))
    return states, index, _sparse.dense_matrix(P_default, states, index)


def _run(sched, states, index, B, node_to_allowed_states, root_distn, edge_times, seed):
    S = len(states)
    full = (1 << S) - 1
    mask = np.full((sched.n, 1), full, dtype=np.uint64)
    if node_to_allowed_states:
        for v, allowed in node_to_allowed_states.items():
            i = sched.node_index.get(v)
            if i is not None:
                m = 0
                for s in allowed:
                    if s in index:
                        m |= 1 << index[s]
                mask[i, 0] = m
    prior = None
    if root_distn is not None:
        prior = np.array([root_distn.get(s, 0.0) for s in states], dtype=float)
    obs = engine.Observations.from_masks(sched, mask)
    n_events = sum(len(v) for v in edge_times.values())
    chain = RaoTehChains(sched, None, obs, n_chains=1, root_distn=prior, cap=max(16, n_events + 4),
                         seed=int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else seed,
                         chain_matrix=B)
    chain.load_events(edge_times)
    chain._call(1, -1, False)          # one FFBS pass over the loaded events; status handled here
    st = int(chain.status[0])
    if st == 1:
        raise StructuralZeroProb('no assignment of states is feasible')
    if st != 0:
        raise _native.NativeError('chain sampler failed with status %d' % st)
    return chain.trajectory(0)


_MAX_EVENTS_PER_BASE_EDGE = 200


def resample_states(T, root, node_to_allowed_states=None, root_distn=None, P_default=None, seed=None):
    """raoteh/sampler/_sample_mcy.py:19-83 -> dict node -> sampled state (one step of the
    transition matrix per edge).  Raises StructuralZeroProb when nothing is feasible."""
    if root not in T:
        raise ValueError('the specified root is not in the tree')
    for a, b in T.edges():
        if T[a][b].get('P', None) is not None:
            raise NotImplementedError('per-edge transition matrices are not supported: pass P_default')
    states, index, B = _states_and_matrix(P_default, root_distn, node_to_allowed_states)
    base = TreeSchedule.from_nx(T, root)
    sched = TreeSchedule(base.parent, np.where(np.arange(base.n) > 0, 1.0, 0.0), base.nodes)
    edge_times = dict((c, [0.5]) for c in range(1, sched.n))
    ns, edges = _run(sched, states, index, B, node_to_allowed_states, root_distn, edge_times, seed)
    return dict((sched.nodes[i], states[int(ns[i])]) for i in range(sched.n))


def resample_edge_states(T, root, P, event_nodes, node_to_allowed_states=None, root_distn=None,
                         seed=None):
    """raoteh/sampler/_sample_mcy.py:86-187 -> copy of T whose edges carry `state`; the state is
    constant across non-event nodes and takes one step of P at every event node
    (the chunk tree of raoteh/sampler/_graph_transform.py:298 is implicit)."""
    if root not in T:
        raise ValueError('the specified root is not in the tree')
    if root in event_nodes:
        raise ValueError('the root cannot be an event node')
    if node_to_allowed_states and set(node_to_allowed_states) & set(event_nodes):
        raise NotImplementedError('state restrictions on event nodes are not supported')
    for v in event_nodes:
        if T.degree(v) != 2:
            raise ValueError('an event node must have degree two')
    states, index, B = _states_and_matrix(P, root_distn, node_to_allowed_states)
    # contract the event nodes: base edges are paths between consecutive non-event nodes
    base_nodes = [root]
    parent, length, times, path_of = [-1], [0.0], {}, {}
    pos = {root: 0}
    stack = [(root, None)]
    while stack:
        a, came = stack.pop()
        for b in T[a]:
            if b == came:
                continue
            prev, cur, t, ev, path, t0 = a, b, 0.0, [], [a], 0.0
            above = pos[a]
            while True:
                w = T[prev][cur].get('weight', None)
                step = 1.0 if not w else float(w)
                if len(ev) >= _MAX_EVENTS_PER_BASE_EDGE:
                    # the kernels count events per branch in uint8: cut the chain of event nodes
                    # in the middle of the tree edge (prev, cur) with an unrestricted degree-two
                    # pseudo node (no event there, so the state is the same on both sides)
                    cut = len(base_nodes)
                    base_nodes.append(('__cut__', cut))
                    parent.append(above)
                    length.append(t + 0.5 * step)
                    times[cut] = ev
                    path_of[cut] = (t0, path)
                    above, t, t0, ev, path = cut, -0.5 * step, -0.5 * step, [], [prev]
                t += step
                path.append(cur)
                if cur not in event_nodes:
                    break
                ev.append(t)
                nxt = [x for x in T[cur] if x != prev][0]
                prev, cur = cur, nxt
            pos[cur] = len(base_nodes)
            base_nodes.append(cur)
            parent.append(above)
            length.append(t)
            times[pos[cur]] = ev
            path_of[pos[cur]] = (t0, path)
            stack.append((cur, prev))
    # preorder check: parents were appended before children by construction
    sched = TreeSchedule(np.asarray(parent, dtype=np.int32), np.asarray(length), base_nodes)
    ns, edges = _run(sched, states, index, B, node_to_allowed_states, root_distn, times, seed)
    T_aug = nx.Graph()
    for c, (t0, path) in path_of.items():
        jump_times, seg_states = edges[c]
        t = t0
        for u, v in zip(path[:-1], path[1:]):
            w = T[u][v].get('weight', None)
            step = 1.0 if not w else float(w)
            mid = t + 0.5 * step
            k = int(np.searchsorted(np.asarray(jump_times), mid))
            attrs = dict(state=states[int(seg_states[k])])
            if w is not None:
                attrs['weight'] = w
            T_aug.add_edge(u, v, **attrs)
            t += step
    return T_aug
