"""
TEST INFRASTRUCTURE -- CPU oracle for the tolerance-process sampler (SURVEY rows
A17/A18).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product (raoteh_b200/) never does.

What the reference's blocked Gibbs sampler targets
--------------------------------------------------
`_sample_tmjp_dense.gen_histories_v1` (raoteh/sampler/_sample_tmjp_dense.py:40-171)
"addresses the dependence among components of the tolerance process strictly through
conditioning rather than through rate dependence" (:187-189, :389-391):

  * resample_primary_states_v1 (:175-371) is a Rao-Teh update of the primary
    trajectory under the UNMODIFIED primary process MJP(Q_primary) (Poisson rates
    omega - q_s with the full exit rates q_s, :78-83; P_primary = I + Q_primary/omega),
    conditioned on the event "the class of the primary state is ON at all times";
  * resample_tolerance_states_v1 (:374-506) is a Rao-Teh update of one tolerance
    class under the UNMODIFIED 2-state process [[-on, on], [off, -off]] (:87-96),
    conditioned on the same event (and on the disease data).

Both are exact conditionals of ONE joint law: the PRODUCT of the independent
processes MJP(Q_primary) x prod_c MJP2(rate_on, rate_off), started from
primary_distn x prod_c tolerance_distn, conditioned on compatibility at all times and
on the data.  Conditioning a product of Markov processes on never leaving a set is a
KILLED Markov process on the compatible compound states:

    (p, t) -> (p', t)   rate Q[p, p']   if t[part(p')] = 1     (else: killed)
    (p, t) -> (p, t^c)  rate_on / rate_off                      (off of class part(p): killed)
    diagonal: -( q_p  +  sum_c (rate_off if t_c else rate_on) )   <- FULL exit rates

so every posterior expectation of the sampler's target is available in closed form from
the same pruning + Frechet machinery as any MJP (np_oracle.expected_history_statistics,
which is pinned to the reference's golden vectors) applied to this sub-generator.
`tests/golden/tmjp_v1_moments.json` (oracle/gen_golden.py gen_tmjp_moments) holds the
moments of the reference's own gen_histories_v1 run in the build container; the test
`tests/test_oracle_golden.py::test_killed_generator_matches_reference_sampler` checks this
closed form against them, which pins the statement above to the reference.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import np_oracle


def compound_space(S, part, n_parts):
    """Compatible compound states: (primary p, tolerance tuple t) with t[part[p]] == 1."""
    states = []
    for p in range(S):
        for t in itertools.product((0, 1), repeat=n_parts):
            if t[part[p]] == 1:
                states.append((p, t))
    return states


def killed_generator(Q_primary, part, n_parts, rate_on, rate_off):
    """Sub-generator of the product law conditioned on compatibility (see module doc)."""
    Q = np.asarray(Q_primary, dtype=float)
    S = Q.shape[0]
    states = compound_space(S, part, n_parts)
    index = dict((s, i) for i, s in enumerate(states))
    n = len(states)
    K = np.zeros((n, n))
    q = -np.diag(Q)
    for i, (p, t) in enumerate(states):
        for p2 in range(S):
            if p2 != p and Q[p, p2] > 0 and t[part[p2]] == 1:
                K[i, index[(p2, t)]] = Q[p, p2]
        for c in range(n_parts):
            t2 = list(t)
            t2[c] = 1 - t[c]
            t2 = tuple(t2)
            if (p, t2) in index:
                K[i, index[(p, t2)]] = rate_on if t2[c] else rate_off
        K[i, i] = -(q[p] + sum(rate_off if tc else rate_on for tc in t))
    return states, K


def compound_root_weights(states, primary_distn, rate_on, rate_off):
    """primary_distn x prod_c tolerance_distn (raoteh/sampler/_tmjp_dense.py:353-377)."""
    tol = np.array([rate_off, rate_on], dtype=float) / (rate_on + rate_off)
    return np.array([primary_distn[p] * np.prod([tol[tc] for tc in t]) for p, t in states])


def expected_sampler_statistics(parent, lengths, Q_primary, part, n_parts, primary_distn,
                                rate_on, rate_off, node_to_primary_state, disease_data=None):
    """Closed-form posterior expectations under the v1 sampler's target law, one site.

    node_to_primary_state: dict node -> observed primary state;
    disease_data: list over classes of dict node -> set of allowed tolerance states.
    Returns dict(prim_dwell[S], prim_trans[S,S], tol[n_parts,4] = (root on, dwell on, gains,
    losses), loglik).
    """
    Q = np.asarray(Q_primary, dtype=float)
    S = Q.shape[0]
    states, K = killed_generator(Q, part, n_parts, rate_on, rate_off)
    n = len(parent)
    nc = len(states)
    mask = np.ones((n, 1, nc), dtype=bool)
    for v in range(n):
        for i, (p, t) in enumerate(states):
            ok = True
            if v in node_to_primary_state and node_to_primary_state[v] != p:
                ok = False
            if disease_data is not None:
                for c in range(n_parts):
                    if v in disease_data[c] and t[c] not in disease_data[c][v]:
                        ok = False
            mask[v, 0, i] = ok
    lik = mask.astype(float)
    obs = np_oracle.Obs('dense', nc, 1, lik=lik, has=np.ones(n, dtype=bool))
    w = compound_root_weights(states, primary_distn, rate_on, rate_off)
    P = np_oracle.expm_edges(K, lengths)
    r = np_oracle.expected_history_statistics(parent, lengths, K, P, obs, w)
    prim_dwell = np.zeros(S)
    prim_trans = np.zeros((S, S))
    tol = np.zeros((n_parts, 4))
    root_post = r['root_post'][0]
    for i, (p, t) in enumerate(states):
        prim_dwell[p] += r['dwell'][i]
        for c in range(n_parts):
            if t[c]:
                tol[c, 0] += root_post[i]
                tol[c, 1] += r['dwell'][i]
        for j, (p2, t2) in enumerate(states):
            if i == j or K[i, j] == 0:
                continue
            if p != p2:
                prim_trans[p, p2] += r['trans'][i, j]
            else:
                c = [k for k in range(n_parts) if t[k] != t2[k]][0]
                tol[c, 2 if t2[c] else 3] += r['trans'][i, j]
    return dict(prim_dwell=prim_dwell, prim_trans=prim_trans, tol=tol, loglik=float(r['loglik'][0]))
