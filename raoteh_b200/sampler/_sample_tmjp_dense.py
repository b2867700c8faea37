"""
Rao-Teh samples of tolerance MJP trajectories on trees, dense model
(raoteh/sampler/_sample_tmjp_dense.py).  `disease_data` is a list, indexed by tolerance
class, of maps from a node to a set of allowed tolerance states.

The generator owns one device-resident compound trajectory (a batch of one chain x one
site of raoteh_b200.tmjp.ToleranceChains); every `next()` runs one blocked Gibbs sweep of
the CUDA kernel (csrc/rt_tmjp.cu) and converts the trajectories to the reference's output
type.  Batched callers use ToleranceChains directly.
"""
from __future__ import division, print_function, absolute_import

import itertools

import networkx as nx
import numpy as np

from .. import engine
from ..lowering import TreeSchedule, check_square_dense
from ..tmjp import ToleranceChains

__all__ = []


def _edges_to_graph(sched, label, node_values, edges, next_node):
    """Per-edge (jump times, segment states) -> nx.Graph with `weight` and `state`;
    new degree-2 node ids start at next_node.  Returns (graph, next free id)."""
    G = nx.Graph()
    for i in range(1, sched.n):
        a, b = sched.nodes[sched.parent[i]], sched.nodes[i]
        times, seg_states = edges[i]
        prev_node, prev_t = a, 0.0
        for j, tau in enumerate(times):
            mid = next_node
            next_node += 1
            G.add_edge(prev_node, mid, weight=float(tau - prev_t), state=label(seg_states[j]))
            prev_node, prev_t = mid, float(tau)
        G.add_edge(prev_node, b, weight=float(sched.length[i] - prev_t), state=label(seg_states[-1]))
    return G, next_node


def _lower_inputs(T, root, nprimary, node_to_primary_state, disease_data, nparts, state_index=None):
    if root not in T:
        raise ValueError('the root must be a node in the tree')
    sched = TreeSchedule.from_nx(T, root)
    obs_nodes = sorted(sched.node_index[v] for v in node_to_primary_state if v in sched.node_index)
    codes = np.full((max(1, len(obs_nodes)), 1), 255, dtype=np.uint8)
    for k, i in enumerate(obs_nodes):
        s = node_to_primary_state[sched.nodes[i]]
        codes[k, 0] = s if state_index is None else state_index[s]
    if not obs_nodes:
        obs_nodes = [int(sched.leaves[0])] if len(sched.leaves) else [0]
    obs_slot = np.full(sched.n, -1, dtype=np.int32)
    obs_slot[obs_nodes] = np.arange(len(obs_nodes), dtype=np.int32)
    import torch
    obs = engine.Observations(engine.OBS_CODES, torch.from_numpy(codes).cuda(), obs_slot, 1)
    tol_obs = tol_nodes = None
    if disease_data is not None:
        nodes = sorted(set(sched.node_index[v] for d in disease_data for v in d if v in sched.node_index))
        if nodes:
            tol_nodes = nodes
            tol_obs = np.full((len(nodes), nparts, 1), 3, dtype=np.uint8)
            for c, d in enumerate(disease_data):
                for v, allowed in d.items():
                    if v in sched.node_index:
                        bits = sum(1 << int(s) for s in allowed if s in (0, 1))
                        tol_obs[nodes.index(sched.node_index[v]), c, 0] = bits
    return sched, obs, tol_obs, tol_nodes


def _gen(sched, chains, primary_label, T, nhistories):
    try:
        chains.initialize()
    except RuntimeError:
        raise Exception('failed to find a feasible history')
    base_next = max(T) + 1
    for i in itertools.count():
        next_node = base_next
        ns, edges = chains.primary_trajectory(0)
        primary, next_node = _edges_to_graph(sched, primary_label, ns, edges, next_node)
        tolerance = []
        for c in range(chains.n_parts):
            bits, tedges = chains.tolerance_trajectory(0, c)
            G, next_node = _edges_to_graph(sched, int, bits, tedges, next_node)
            tolerance.append(G)
        yield primary, tolerance
        if nhistories is not None and i + 1 >= nhistories:
            return
        chains.sweep(1, stats=False)


def gen_histories_v1(ctm, T, root, node_to_primary_state, disease_data=None,
                     uniformization_factor=2, nhistories=None, seed=None, cap_p=None, cap_t=None):
    """raoteh/sampler/_sample_tmjp_dense.py:40-171: generator of
    (primary_trajectory, [tolerance_trajectory] * nparts), each an nx.Graph whose edges
    carry `weight` and `state`; redundant degree-2 nodes are already removed."""
    check_square_dense(ctm.Q_primary)
    sched, obs, tol_obs, tol_nodes = _lower_inputs(
        T, root, ctm.nprimary, node_to_primary_state, disease_data, ctm.nparts)
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    Q = np.asarray(ctm.Q_primary, dtype=float)
    total = float(sched.length.sum())
    if cap_p is None:
        cap_p = int(min(4096, max(64, 6 * uniformization_factor * np.max(-np.diag(Q)) * total + 4 * sched.n)))
    if cap_t is None:
        cap_t = int(min(255, max(48, 6 * uniformization_factor * max(ctm.rate_on, ctm.rate_off) * total + 32)))
    chains = ToleranceChains(sched, Q, np.asarray(ctm.primary_distn, dtype=float), ctm.primary_to_part,
                             ctm.rate_on, ctm.rate_off, obs, n_chains=1, tol_obs=tol_obs,
                             tol_obs_nodes=tol_nodes, uniformization_factor=uniformization_factor,
                             cap_p=cap_p, cap_t=cap_t, seed=seed)
    for history in _gen(sched, chains, int, T, nhistories):
        yield history
