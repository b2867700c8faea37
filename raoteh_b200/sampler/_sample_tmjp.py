"""
Rao-Teh samples of tolerance MJP trajectories on trees, sparse model
(raoteh/sampler/_sample_tmjp.py:34-168): the primary rate matrix is a weighted nx.DiGraph
without diagonal, distributions are dicts, primary states are arbitrary sortable labels.
Lowered to the dense device sampler of _sample_tmjp_dense.
"""
from __future__ import division, print_function, absolute_import

import numpy as np

from . import _sparse
from ._sample_tmjp_dense import _lower_inputs, _gen
from ..tmjp import ToleranceChains

__all__ = []


def gen_histories(ctm, T, root, node_to_primary_state, disease_data=None,
                  uniformization_factor=2, nhistories=None, seed=None, cap_p=None, cap_t=None):
    """raoteh/sampler/_sample_tmjp.py:34-168: generator of
    (primary_trajectory, [tolerance_trajectory] * nparts)."""
    states = sorted(ctm.primary_to_part)
    index = dict((s, i) for i, s in enumerate(states))
    Q = _sparse.dense_matrix(ctm.Q_primary, states, index)
    Q -= np.diag(Q.sum(axis=1))
    distn = np.array([ctm.primary_distn.get(s, 0.0) for s in states], dtype=float)
    part = dict((index[s], c) for s, c in ctm.primary_to_part.items())
    sched, obs, tol_obs, tol_nodes = _lower_inputs(
        T, root, len(states), node_to_primary_state, disease_data, ctm.nparts, state_index=index)
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    total = float(sched.length.sum())
    if cap_p is None:
        cap_p = int(min(4096, max(64, 6 * uniformization_factor * np.max(-np.diag(Q)) * total + 4 * sched.n)))
    if cap_t is None:
        cap_t = int(min(255, max(48, 6 * uniformization_factor * max(ctm.rate_on, ctm.rate_off) * total + 32)))
    chains = ToleranceChains(sched, Q, distn, part, ctm.rate_on, ctm.rate_off, obs, n_chains=1,
                             tol_obs=tol_obs, tol_obs_nodes=tol_nodes,
                             uniformization_factor=uniformization_factor, cap_p=cap_p, cap_t=cap_t,
                             seed=seed)
    for history in _gen(sched, chains, lambda i: states[i], T, nhistories):
        yield history
