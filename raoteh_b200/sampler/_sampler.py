"""
Rao-Teh history generators with the reference's signatures
(raoteh/sampler/_sampler.py:238-390).  Each generator owns one device-resident
trajectory (a batch of one chain x one site of raoteh_b200.raoteh.RaoTehChains);
every `next()` runs one sweep of the CUDA kernel and converts the trajectory to
the reference's output type: an undirected nx.Graph whose edges carry `weight`
(segment length) and `state`, original node ids preserved, new degree-2 node ids
> max(T), total weight preserved.

Randomness comes from the counter-based Philox generator of the kernel, seeded
from numpy's global RNG at generator creation (the reference draws from the
global `np.random` / `random` state and is never seeded by the library).
"""
from __future__ import division, print_function, absolute_import

import itertools

import networkx as nx
import numpy as np

from . import _mjp, _sparse
from ._util import get_first_element
from .. import engine
from ..lowering import TreeSchedule, subdivide_long_branches
from ..raoteh import RaoTehChains

__all__ = []


def _merged_trajectory(chain):
    """Trajectory 0 of the chain on the ORIGINAL tree: pieces of subdivided branches joined."""
    ns, edges = chain.trajectory(0)
    sched0, pieces = chain.merge
    ns0 = np.zeros(sched0.n, dtype=int)
    ns0[0] = ns[0]
    edges0 = {}
    for c in range(1, sched0.n):
        times, seg_states, offset = [], None, 0.0
        for node in pieces[c]:
            tt, ss = edges[node]
            if seg_states is None:
                seg_states = [int(ss[0])]
            for tau, s_after in zip(tt, ss[1:]):
                times.append(offset + float(tau))
                seg_states.append(int(s_after))
            offset += float(chain.sched.length[node])
        ns0[c] = seg_states[-1]
        edges0[c] = (np.asarray(times, dtype=float), np.asarray(seg_states, dtype=int))
    return ns0, edges0


def _trajectory_to_graph(sched, states, ns, edges, next_node):
    T_out = nx.Graph()
    for i in range(1, sched.n):
        a, b = sched.nodes[sched.parent[i]], sched.nodes[i]
        times, seg_states = edges[i]
        prev_node, prev_t = a, 0.0
        for j, tau in enumerate(times):
            mid = next_node
            next_node += 1
            T_out.add_edge(prev_node, mid, weight=float(tau - prev_t), state=states[seg_states[j]])
            prev_node, prev_t = mid, float(tau)
        # float32 event times can exceed the fp64 branch length by one rounding step
        T_out.add_edge(prev_node, b, weight=max(0.0, float(sched.length[i] - prev_t)),
                       state=states[seg_states[-1]])
    return T_out


def _make_chain(T, Q, node_to_allowed_states, root, root_distn, uniformization_factor, seed, cap):
    """Validation and lowering shared by the history generators
    (raoteh/sampler/_sampler.py:329-362, :431-470)."""
    bad = set(node_to_allowed_states) - set(T)
    if bad:
        raise ValueError('some of the nodes which have been annotated with state restrictions '
                         'are not even in the tree: ' + str(sorted(bad)))
    if uniformization_factor <= 1:
        raise ValueError('the uniformization factor must be greater than 1')
    if not Q:
        raise ValueError('the rate matrix is empty')
    for a, b in Q.edges():
        if a == b:
            raise ValueError('the rate matrix should have no loops')
    if root not in T:
        raise ValueError('the root must be a node in the tree')
    states = sorted(set(Q) | (set(root_distn) if root_distn is not None else set()))
    index = dict((s, i) for i, s in enumerate(states))
    S = len(states)
    Qd = _sparse.dense_matrix(Q, states, index)
    Qd -= np.diag(Qd.sum(axis=1))
    sched0 = TreeSchedule.from_nx(T, root)
    # long branches are cut into pieces (new unobserved degree-two nodes) so that no branch
    # expects more than ~100 candidate events per sweep; the pieces are merged again when a
    # history is converted to the reference's graph type
    omega = uniformization_factor * np.max(-np.diag(Qd))
    sched, pieces, image = subdivide_long_branches(sched0, 100.0 / omega if omega > 0 else np.inf)
    full = (1 << S) - 1
    mask = np.full((sched.n, 1), full, dtype=np.uint64)
    for v, allowed in node_to_allowed_states.items():
        m = 0
        for s in allowed:
            if s in index:
                m |= 1 << index[s]
        mask[image[sched0.node_index[v]], 0] = m
    prior = None
    if root_distn is not None:
        prior = np.array([root_distn.get(s, 0.0) for s in states], dtype=float)
    obs = engine.Observations.from_masks(sched, mask)
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    if cap is None:
        mean = omega * sched.length.sum()
        cap = int(max(256, (sched.n - 1) * (S + 1), 4 * mean + 64))
    chain = RaoTehChains(sched, Qd, obs, n_chains=1, root_distn=prior,
                         uniformization_factor=uniformization_factor, cap=cap, seed=seed)
    try:
        chain.initialize()
    except RuntimeError:
        raise Exception('failed to find a feasible history')
    chain.merge = (sched0, pieces)
    return chain, sched0, states


def gen_restricted_histories(T, Q, node_to_allowed_states, root, root_distn=None,
                             uniformization_factor=2, nhistories=None, seed=None, cap=None):
    """raoteh/sampler/_sampler.py:300-390 (generator of nx.Graph histories)."""
    chain, sched, states = _make_chain(T, Q, node_to_allowed_states, root, root_distn,
                                       uniformization_factor, seed, cap)
    next_node = max(T) + 1
    for i in itertools.count():
        ns, edges = _merged_trajectory(chain)
        yield _trajectory_to_graph(sched, states, ns, edges, next_node)
        if nhistories is not None and i + 1 >= nhistories:
            return
        chain.sweep(1, stats=False)
        chain.check()


def gen_mh_histories(T, Q, node_to_allowed_states, target_log_likelihood_callback,
                     root, root_distn=None, uniformization_factor=2, nhistories=None,
                     seed=None, cap=None):
    """raoteh/sampler/_sampler.py:393-551: Rao-Teh sweeps as Metropolis-Hastings proposals.
    Yields (history, accept_flag); a rejected proposal re-yields the previous history.  The
    proposal density is the trajectory likelihood under Q (_mjp.get_trajectory_log_likelihood,
    :462-463), the target is the caller's callback on the history graph; the accept draw
    uses Python's `random` like the reference (:527)."""
    import random
    chain, sched, states = _make_chain(T, Q, node_to_allowed_states, root, root_distn,
                                       uniformization_factor, seed, cap)
    next_node = max(T) + 1

    def biased(T_aug):
        return _mjp.get_trajectory_log_likelihood(T_aug, root, root_distn, Q)

    T_prev = None
    saved = None
    ll_biased_prev = ll_target_prev = None
    for i in itertools.count():
        ns, edges = _merged_trajectory(chain)
        T_cur = _trajectory_to_graph(sched, states, ns, edges, next_node)
        if T_prev is None:
            accept_flag = True
        else:
            if ll_biased_prev is None:
                ll_biased_prev = biased(T_prev)
            ll_biased_curr = biased(T_cur)
            if ll_target_prev is None:
                ll_target_prev = target_log_likelihood_callback(T_prev)
            ll_target_curr = target_log_likelihood_callback(T_cur)
            log_mh_ratio = ll_target_curr - ll_target_prev - ll_biased_curr + ll_biased_prev
            accept_flag = bool(log_mh_ratio > 0 or random.random() < np.exp(log_mh_ratio))
            if accept_flag:
                ll_biased_prev, ll_target_prev = ll_biased_curr, ll_target_curr
        if not accept_flag:
            T_cur = T_prev
            chain.restore(saved)          # the device trajectory goes back to the previous sample
        yield T_cur, accept_flag
        T_prev = T_cur
        if nhistories is not None and i + 1 >= nhistories:
            return
        saved = chain.snapshot()
        chain.sweep(1, stats=False)
        chain.check()


def gen_histories(T, Q, node_to_state, root=None, root_distn=None, uniformization_factor=2,
                  nhistories=None, seed=None):
    """raoteh/sampler/_sampler.py:238-297"""
    if root is None:
        root = get_first_element(node_to_state) if node_to_state else get_first_element(T)
    all_states = set(Q)
    if root_distn is not None:
        all_states.update(set(root_distn))
    node_to_allowed_states = {}
    for node in T:
        if node in node_to_state:
            node_to_allowed_states[node] = {node_to_state[node]}
        else:
            node_to_allowed_states[node] = all_states
    for history in gen_restricted_histories(T, Q, node_to_allowed_states, root,
                                            root_distn=root_distn,
                                            uniformization_factor=uniformization_factor,
                                            nhistories=nhistories, seed=seed):
        yield history


def get_restricted_feasible_history(T, P, node_to_allowed_states, root, root_distn=None):
    """raoteh/sampler/_sampler.py:563-643: an arbitrary feasible history under the
    uniformized matrix P (weighted DiGraph with self-loops); the law is not meaningful."""
    states = sorted(P)
    index = dict((s, i) for i, s in enumerate(states))
    B = _sparse.dense_matrix(P, states, index)
    Q = nx.DiGraph()
    for a, b in P.edges():
        if a != b and P[a][b]['weight'] > 0:
            Q.add_edge(a, b, weight=P[a][b]['weight'])
    if not Q:
        raise Exception('failed to find a feasible history')
    gen = gen_restricted_histories(T, Q, node_to_allowed_states, root, root_distn=root_distn,
                                   nhistories=1)
    return next(gen)


def get_forward_sample(T, Q, root, root_distn):
    """raoteh/sampler/_sampler.py:163-235: unconditional forward simulation of one history
    (host-side data generation, jump by jump with numpy's global RNG like the reference; it is
    not on the accelerated path).  Returns the augmented tree with `state` and `weight`."""
    total_rates = _mjp.get_total_rates(Q)
    P = _mjp.get_conditional_transition_matrix(Q, total_rates)
    next_node = max(T) + 1
    states, probs = zip(*root_distn.items())
    node_to_state = {root: states[int(np.random.choice(len(states), p=probs))]}
    T_out = nx.Graph()
    for a, b in nx.bfs_edges(T, root):
        state = node_to_state[a]
        weight = T[a][b]['weight']
        prev_node, total_dwell = a, 0.0
        while state in total_rates and total_rates[state] > 0:
            dwell = np.random.exponential(scale=1.0 / total_rates[state])
            if total_dwell + dwell > weight:
                break
            total_dwell += dwell
            mid_node = next_node
            next_node += 1
            T_out.add_edge(prev_node, mid_node, state=state, weight=dwell)
            prev_node = mid_node
            nxt = list(P[state])
            state = nxt[int(np.random.choice(len(nxt), p=[P[state][s]['weight'] for s in nxt]))]
            node_to_state[prev_node] = state
        node_to_state[b] = state
        T_out.add_edge(prev_node, b, state=state, weight=weight - total_dwell)
    return T_out


def gen_forward_samples(T, Q, root, root_distn, nsamples=None):
    """raoteh/sampler/_sampler.py:67-160: unconditional forward samples, one augmented tree per
    draw (host code like the reference: it makes test data, it is not on the accelerated path)."""
    for i in itertools.count():
        if i == nsamples:
            return
        yield get_forward_sample(T, Q, root, root_distn)
