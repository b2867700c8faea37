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


# ---------------------------------------------------------------------------------------
# CPU restatement of one blocked Gibbs sweep (one trajectory at a time; small cases and the
# CPU leg of bench.py).  Follows raoteh/sampler/_sample_tmjp_dense.py:116-171:
#   resample_poisson on the primary trajectory (_sample_mjp_dense.py:21-69),
#   resample_primary_states_v1 (:175-371), then per class resample_poisson (_sample_mjp.py:19)
#   and resample_tolerance_states_v1 (:374-506).
# A trajectory is {'node': states at the tree nodes, 'edges': {child: (jump times from the
# parent end, ascending; states of the k+1 segments, parent side first)}}.
# ---------------------------------------------------------------------------------------
def _pieces(traj, c, length):
    """[(t_lo, t_hi, state)] of edge c from the parent end to the child end."""
    times, states = traj['edges'][c]
    b = [0.0] + list(times) + [length]
    return [(b[i], b[i + 1], states[i]) for i in range(len(states))]


def _poisson_events(pieces, rates, rng):
    """Fresh events of a Poisson process of rate rates[state] on every piece."""
    out = []
    for lo, hi, s in pieces:
        r = rates[s]
        if r <= 0 or hi <= lo:
            continue
        k = rng.poisson(r * (hi - lo))
        out.extend(rng.uniform(lo, hi, size=k).tolist())
    return out


def _ffbs_tree(parent, lengths, events, chunk_weight, node_weight, B, root_w, rng):
    """FFBS over the implicit chunk tree.  events[c]: sorted candidate times on edge c;
    chunk_weight(c, lo, hi) -> state weights (0/1) of the piece (lo, hi) of edge c;
    node_weight(v) -> state weights at node v.  Returns (node states, per-edge states of the
    len(events[c]) + 1 pieces, parent side first)."""
    n = len(parent)
    S = B.shape[0]
    children = [[] for _ in range(n)]
    for b in range(1, n):
        children[parent[b]].append(b)
    msgs = {}
    beta_below = {}
    partial = [None] * n
    for v in range(n - 1, -1, -1):
        acc = node_weight(v).astype(float).copy()
        for c in children[v]:
            acc *= msgs[c]
        partial[v] = acc
        if v == 0:
            break
        ev = events[v]
        bounds = [0.0] + list(ev) + [lengths[v]]
        beta = partial[v].copy()
        rec = []
        for i in range(len(ev), -1, -1):          # pieces from the child end upwards
            beta = beta * chunk_weight(v, bounds[i], bounds[i + 1])
            if i > 0:
                rec.append(beta.copy())            # message just below event i-1
                beta = B @ beta
                m = beta.max()
                if m > 0:
                    beta = beta / m
        msgs[v] = beta
        beta_below[v] = rec[::-1]                  # indexed by event, parent side first
    w = root_w * partial[0]
    if not w.sum() > 0:
        raise ValueError('infeasible')
    node = np.zeros(n, dtype=int)
    node[0] = rng.choice(S, p=w / w.sum())
    seg_states = {}
    for v in range(1, n):
        cur = node[parent[v]]
        st = [cur]
        for j in range(len(events[v])):
            w = B[cur] * beta_below[v][j]
            cur = rng.choice(S, p=w / w.sum())
            st.append(cur)
        seg_states[v] = st
        node[v] = cur
    return node, seg_states


def _compress(events, seg_states, node, n):
    """Drop self-transitions."""
    edges = {}
    for c in range(1, n):
        times, states = [], [seg_states[c][0]]
        for t, s in zip(events[c], seg_states[c][1:]):
            if s != states[-1]:
                times.append(t)
                states.append(s)
        edges[c] = (times, states)
    return dict(node=node, edges=edges)


def gibbs_sweep(parent, lengths, Q_primary, part, n_parts, primary_distn, rate_on, rate_off,
                node_to_primary_state, disease_data, prim, tols, rng, uniformization_factor=2.0):
    """One sweep: primary given all tolerance trajectories, then every class given the primary."""
    Q = np.asarray(Q_primary, dtype=float)
    S = Q.shape[0]
    n = len(parent)
    part = np.asarray(part)
    q = -np.diag(Q)
    omega = uniformization_factor * q.max()
    B = np.eye(S) + Q / omega
    # ---- primary (resample_primary_states_v1)
    events = {}
    for c in range(1, n):
        pcs = _pieces(prim, c, lengths[c])
        events[c] = sorted(list(prim['edges'][c][0]) + _poisson_events(pcs, omega - q, rng))
    tol_pieces = [[None] + [_pieces(tols[k], c, lengths[c]) for c in range(1, n)] for k in range(n_parts)]

    def prim_chunk(c, lo, hi):
        w = np.ones(S)
        for k in range(n_parts):
            if any(s == 0 and min(hi, b) - max(lo, a) >= 0 and not (b < lo or a > hi)
                   for a, b, s in tol_pieces[k][c]):
                w[part == k] = 0.0
        return w

    def prim_node(v):
        w = np.ones(S)
        if v in node_to_primary_state:
            w[:] = 0.0
            w[node_to_primary_state[v]] = 1.0
        return w
    node, seg = _ffbs_tree(parent, lengths, events, prim_chunk, prim_node, B,
                           np.asarray(primary_distn, dtype=float), rng)
    prim = _compress(events, seg, node, n)
    # ---- tolerance classes (resample_tolerance_states_v1)
    Qt = np.array([[-rate_on, rate_on], [rate_off, -rate_off]])
    omega_t = uniformization_factor * max(rate_on, rate_off)
    Bt = np.eye(2) + Qt / omega_t
    rates_t = omega_t - np.array([rate_on, rate_off])
    distn_t = np.array([rate_off, rate_on]) / (rate_on + rate_off)
    prim_pieces = [None] + [_pieces(prim, c, lengths[c]) for c in range(1, n)]
    new_tols = []
    for k in range(n_parts):
        ev = {}
        for c in range(1, n):
            pcs = _pieces(tols[k], c, lengths[c])
            ev[c] = sorted(list(tols[k]['edges'][c][0]) + _poisson_events(pcs, rates_t, rng))

        def tol_chunk(c, lo, hi, k=k):
            need_on = any(part[s] == k and not (b < lo or a > hi) for a, b, s in prim_pieces[c])
            return np.array([0.0 if need_on else 1.0, 1.0])

        def tol_node(v, k=k):
            w = np.ones(2)
            if disease_data is not None and v in disease_data[k]:
                w = np.array([1.0 if 0 in disease_data[k][v] else 0.0,
                              1.0 if 1 in disease_data[k][v] else 0.0])
            return w
        node_t, seg_t = _ffbs_tree(parent, lengths, ev, tol_chunk, tol_node, Bt, distn_t, rng)
        new_tols.append(_compress(ev, seg_t, node_t, n))
    return prim, new_tols


def gibbs_init(parent, lengths, Q_primary, part, n_parts, primary_distn, rate_on, rate_off,
               node_to_primary_state, disease_data, rng, max_events=7):
    """An arbitrary jointly feasible history (raoteh/sampler/_sample_tmjp_dense.py:509-627):
    primary from equally spaced events under the uniformized matrix, then every class with one
    event at a uniform time inside each primary segment."""
    Q = np.asarray(Q_primary, dtype=float)
    S = Q.shape[0]
    n = len(parent)
    part = np.asarray(part)
    B = np.eye(S) + Q / (2.0 * (-np.diag(Q)).max())

    def prim_node(v):
        w = np.ones(S)
        if v in node_to_primary_state:
            w[:] = 0.0
            w[node_to_primary_state[v]] = 1.0
        return w
    k = 0
    while True:
        events = dict((c, [lengths[c] * (i + 1) / (k + 1) for i in range(k)]) for c in range(1, n))
        try:
            node, seg = _ffbs_tree(parent, lengths, events, lambda c, lo, hi: np.ones(S), prim_node, B,
                                   np.asarray(primary_distn, dtype=float), rng)
            break
        except ValueError:
            k = 2 * k + 1
            if k > max_events:
                raise
    prim = _compress(events, seg, node, n)
    Bt = np.eye(2) + np.array([[-rate_on, rate_on], [rate_off, -rate_off]]) / (2.0 * max(rate_on, rate_off))
    distn_t = np.array([rate_off, rate_on]) / (rate_on + rate_off)
    prim_pieces = [None] + [_pieces(prim, c, lengths[c]) for c in range(1, n)]
    tols = []
    for kk in range(n_parts):
        ev = dict((c, [rng.uniform(a, b) for a, b, s in prim_pieces[c]]) for c in range(1, n))

        def tol_chunk(c, lo, hi, kk=kk):
            need_on = any(part[s] == kk and not (b < lo or a > hi) for a, b, s in prim_pieces[c])
            return np.array([0.0 if need_on else 1.0, 1.0])

        def tol_node(v, kk=kk):
            w = np.ones(2)
            if disease_data is not None and v in disease_data[kk]:
                w = np.array([1.0 if 0 in disease_data[kk][v] else 0.0,
                              1.0 if 1 in disease_data[kk][v] else 0.0])
            return w
        node_t, seg_t = _ffbs_tree(parent, lengths, ev, tol_chunk, tol_node, Bt, distn_t, rng)
        tols.append(_compress(ev, seg_t, node_t, n))
    return prim, tols


def sampled_statistics(parent, lengths, prim, tols, S, n_parts):
    """The statistics the GPU sampler accumulates: primary dwell[S], transitions[S,S], per class
    (root on, dwell on, gains, losses)."""
    n = len(parent)
    dwell = np.zeros(S)
    trans = np.zeros((S, S))
    for c in range(1, n):
        pcs = _pieces(prim, c, lengths[c])
        for a, b, s in pcs:
            dwell[s] += b - a
        for (a0, b0, s0), (a1, b1, s1) in zip(pcs[:-1], pcs[1:]):
            trans[s0, s1] += 1
    tol = np.zeros((n_parts, 4))
    for k in range(n_parts):
        tol[k, 0] = tols[k]['node'][0]
        for c in range(1, n):
            pcs = _pieces(tols[k], c, lengths[c])
            for a, b, s in pcs:
                if s:
                    tol[k, 1] += b - a
            for (a0, b0, s0), (a1, b1, s1) in zip(pcs[:-1], pcs[1:]):
                tol[k, 2 if s1 else 3] += 1
    return dwell, trans, tol
