"""Warp-per-trajectory kernels (csrc/rt_tmjp.cu) on the GPU, through the C ABI (rt_tmjp_run,
rt_raoteh_sweeps for S > 8):

* K7  tolerance summary == the reference's _tmjp_dense.get_tolerance_summary on the golden
      fixture (tests/golden/tolerance_summary.json), relative 1e-9;
* K6w Rao-Teh sweeps for 9..64 states vs closed-form posterior expectations, |z| < 5;
* K8  blocked Gibbs sampler of the compound tolerance process: structural invariants of
      every sampled history and distributional agreement (|z| < 5) with the closed form of
      the sampler's target law (oracle/np_tmjp.py), which tests/test_oracle_golden.py pins
      to the moments of the reference's own gen_histories_v1.
"""
import numpy as np
import pytest

from helpers import load_golden
from oracle import np_oracle, np_tmjp

pytestmark = pytest.mark.gpu

TOY_PRE = np.array([[0, 1, 1, 0, 0, 0], [1, 0, 0, 1, 0, 0], [1, 0, 0, 1, 1, 0],
                    [0, 1, 1, 0, 0, 1], [0, 0, 1, 0, 0, 1], [0, 0, 0, 1, 1, 0]], dtype=float)


def toy_model():
    Q = TOY_PRE - np.diag(TOY_PRE.sum(axis=1))
    Q /= -np.dot(np.ones(6) / 6, np.diag(Q))
    part = np.array([0, 0, 1, 1, 2, 2])
    return Q, np.ones(6) / 6, part


def toy_tree():
    from raoteh_b200.lowering import TreeSchedule
    # nodes 0..5 in preorder: 0-1, 1-2, 2-3, 2-4, 1-5
    parent = np.array([-1, 0, 1, 2, 2, 1], dtype=np.int32)
    length = np.array([0.0, 0.5, 0.7, 0.4, 0.9, 0.6])
    return TreeSchedule(parent, length)


def _chains_for_fixture(sched, Q, part, rate_on, rate_off, n_traj, disease):
    from raoteh_b200 import engine
    from raoteh_b200.tmjp import ToleranceChains
    codes = np.full((1, n_traj), 255, dtype=np.uint8)
    obs = engine.Observations.from_leaf_codes(sched, codes, leaf_nodes=sched.leaves[:1])
    tol_obs = tol_nodes = None
    if disease is not None:
        nodes = sorted(set(int(n) for d in disease for n in d))
        tol_nodes = [sched.node_index[n] for n in nodes]
        tol_obs = np.full((len(nodes), 3, n_traj), 3, dtype=np.uint8)
        for c, d in enumerate(disease):
            for n, allowed in d.items():
                bits = sum(1 << int(s) for s in allowed)
                tol_obs[nodes.index(int(n)), c, :] = bits
    return ToleranceChains(sched, Q, np.ones(6) / 6, dict(enumerate(part)), rate_on, rate_off, obs,
                           n_chains=1, tol_obs=tol_obs, tol_obs_nodes=tol_nodes, cap_p=64, cap_t=16)


def _fixture_trajectory(case, sched):
    """edge list (na, nb, weight, state) of an augmented primary trajectory -> (node states,
    dict child -> (jump times from the parent end, parent-side states))."""
    base = set(sched.nodes)
    succ = {}
    for a, b, w, s in case['edges']:
        succ.setdefault(a, []).append((b, w, s))
    ns = np.zeros(sched.n, dtype=np.uint8)
    jumps = {}
    root = case['root']
    ns[sched.node_index[root]] = succ[root][0][2]
    stack = [root]
    while stack:
        a = stack.pop()
        for b, w, s in succ.get(a, []):
            # follow the chain of degree-2 extra nodes down to the next base node
            times, states, t, cur_state = [], [], w, s
            node = b
            while node not in base:
                (nb2, w2, s2), = succ[node]
                if s2 != cur_state:
                    times.append(t)
                    states.append(cur_state)
                    cur_state = s2
                t += w2
                node = nb2
            c = sched.node_index[node]
            np.testing.assert_allclose(t, sched.length[c], rtol=1e-9)
            ns[c] = cur_state
            jumps[c] = (times, states)
            stack.append(node)
    return ns, jumps


def test_tolerance_summary_kernel_matches_reference_fixture():
    g = load_golden('tolerance_summary.json')
    Q = np.array(g['Q_primary'])
    part = np.array([g['primary_to_part'][str(i)] for i in range(6)])
    sched = toy_tree()
    groups = {}
    for case in g['cases']:
        key = (case['rate_on'], case['rate_off'], repr(case['disease']))
        groups.setdefault(key, []).append(case)
    assert len(groups) >= 2
    for (rate_on, rate_off, _), cases in groups.items():
        ch = _chains_for_fixture(sched, Q, part, rate_on, rate_off, len(cases), cases[0]['disease'])
        trajs = [_fixture_trajectory(c, sched) for c in cases]
        ch.load_primary_trajectories(np.array([t[0] for t in trajs]), [t[1] for t in trajs])
        out = ch.tolerance_summary().cpu().numpy()
        want = np.array([c['out'] for c in cases])
        # caller-loaded trajectories keep their fp64 times for the summary: relative 1e-9
        np.testing.assert_allclose(out, want, rtol=1e-9, atol=1e-12)
        if cases[0]['disease'] is None:
            # compound log-likelihood with the tolerance histories integrated out
            # (_tmjp.get_tolerance_process_log_likelihood) and the plain trajectory log-likelihood
            # (_mjp.get_trajectory_log_likelihood), both from the reference
            got = ch.tolerance_log_likelihood().cpu().numpy()
            np.testing.assert_allclose(got, [c['tol_ll'] for c in cases], rtol=1e-9)
            got = ch.trajectory_log_likelihood().cpu().numpy()
            np.testing.assert_allclose(got, [c['traj_ll'] for c in cases], rtol=1e-9)


def test_tolerance_summary_kernel_matches_generic_path_fp64_times():
    """Same quantity with jump times exactly representable in float32: relative 1e-9 against
    the generic S = 3 per-edge-Q path of the oracle (np_oracle, pinned to the reference)."""
    from raoteh_b200.sampler import _tmjp_dense  # noqa: F401  (mirror importable)
    Q, pi, part = toy_model()
    sched = toy_tree()
    rng = np.random.default_rng(5)
    n_traj = 12
    ns = np.zeros((n_traj, sched.n), dtype=np.uint8)
    jumps = []
    # random trajectories on the toy tree with dyadic jump times
    length = np.array([0.0, 0.5, 0.75, 0.375, 0.875, 0.625])
    from raoteh_b200.lowering import TreeSchedule
    sched = TreeSchedule(sched.parent, length)
    nbrs = [np.nonzero(TOY_PRE[s])[0] for s in range(6)]
    for t in range(n_traj):
        ns[t, 0] = rng.integers(6)
        ej = {}
        for c in range(1, sched.n):
            k = rng.integers(0, 4)
            times = np.sort(rng.choice(np.arange(1, int(length[c] * 64)), size=k, replace=False)) / 64.0
            cur = ns[t, sched.parent[c]]
            sbs = []
            for _ in range(k):
                sbs.append(cur)
                cur = rng.choice(nbrs[cur])
            ns[t, c] = cur
            ej[c] = (list(times), sbs)
        jumps.append(ej)
    # the last two: w = 0 for every segment, and rate_on EQUAL to one of the absorption rates
    # (the reference's 'defective' case, raoteh/sampler/_linalg.py:116-118)
    from raoteh_b200.tmjp import absorption_rates
    a_eq = float(absorption_rates(Q, part, 3)[0, 1])
    assert a_eq > 0
    for rate_on, rate_off in ((1.0, 1.0), (0.3, 2.0), (2.0, 0.0), (a_eq, 0.0)):
        ch = _chains_for_fixture(sched, Q, part, rate_on, rate_off, n_traj, None)
        ch.load_primary_trajectories(ns, jumps)
        out = ch.tolerance_summary().cpu().numpy()
        for t in range(n_traj):
            want = _oracle_summary(sched, Q, part, rate_on, rate_off, ns[t], jumps[t])
            np.testing.assert_allclose(out[t], want, rtol=1e-9, atol=1e-12)


def _oracle_summary(sched, Q, part, rate_on, rate_off, ns, jumps):
    """get_tolerance_summary restated on the oracle's generic path: per class a 3-state
    inhomogeneous MJP on the augmented tree (raoteh/sampler/_tmjp_dense.py:724-855, :965-1078)."""
    n_parts = int(part.max()) + 1
    absorb = np.zeros((6, n_parts))
    off = Q - np.diag(np.diag(Q))
    for c in range(n_parts):
        absorb[:, c] = off[:, part == c].sum(axis=1)
    # augmented tree: base nodes then jump nodes, in preorder
    parent, length, seg_state = [-1], [0.0], [0]
    index = {0: 0}

    def add(par, t, s):
        parent.append(par)
        length.append(t)
        seg_state.append(s)
        return len(parent) - 1
    for c in range(1, sched.n):
        times, sbs = jumps[c]
        cur = index[int(sched.parent[c])]
        prev = 0.0
        for tau, sb in zip(times, sbs):
            cur = add(cur, tau - prev, sb)
            prev = tau
        index[c] = add(cur, sched.length[c] - prev, ns[c])
    parent = np.array(parent)
    length = np.array(length)
    n = len(parent)
    total = length.sum()
    out = np.zeros(7)
    distn = np.array([rate_off, rate_on, 0.0]) / (rate_on + rate_off)
    for c in range(n_parts):
        Qs = np.zeros((n, 3, 3))
        allowed = np.ones((n, 1, 3), dtype=bool)
        allowed[:, :, 2] = False
        for b in range(1, n):
            s = seg_state[b]
            same = part[s] == c
            w = 0.0 if same else rate_off
            r = absorb[s, c]
            Qs[b] = [[-rate_on, rate_on, 0], [w, -w - r, r], [0, 0, 0]]
            if same:
                allowed[b, :, 0] = False
                allowed[parent[b], :, 0] = False
        P = np_oracle.expm_edges(Qs, length, q_index=np.arange(n))
        obs = np_oracle.Obs('dense', 3, 1, lik=allowed.astype(float), has=np.ones(n, dtype=bool))
        r = np_oracle.expected_history_statistics(parent, length, Qs, P, obs, distn,
                                                  q_index=np.arange(n))
        M = r['M_edges']
        for b in range(1, n):
            out[2] += M[b, 1, 1]
            out[5] += Qs[b, 0, 1] * M[b, 0, 1]
            out[6] += Qs[b, 1, 0] * M[b, 1, 0]
            out[4] += Qs[b, 1, 2] * M[b, 1, 1]
        out[0] += r['root_post'][0][1]
    out[1] = n_parts - out[0]
    out[3] = total * n_parts - out[2]
    return out


# ---------------------------------------------------------------------------------------
# K6w: plain Rao-Teh for 9..64 states
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize('S,n_leaves', [(11, 6), (20, 5), (40, 4)])
def test_large_state_sweeps_match_closed_form(S, n_leaves):
    from raoteh_b200 import synth, engine
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
    rng = np.random.default_rng(300 + S)
    parent, length, leaves = synth.random_binary_tree(n_leaves, 0.25, rng)
    Q = rng.exponential(1.0, size=(S, S)) * (rng.random((S, S)) < 0.4)
    Q += np.roll(np.eye(S), 1, axis=1) * 0.3      # keep it irreducible
    np.fill_diagonal(Q, 0)
    Q -= np.diag(Q.sum(axis=1))
    Q /= np.abs(np.diag(Q)).mean()
    pi = rng.dirichlet(np.ones(S) * 3)
    n_sites = 2
    codes = synth.simulate_leaf_codes(parent, length, leaves, Q, pi, n_sites, rng, 0.1)
    sched = TreeSchedule(parent, length)
    obs = engine.Observations.from_leaf_codes(sched, codes, leaves)
    P = np_oracle.expm_edges(Q, length)
    o = np_oracle.expected_history_statistics(
        parent, length, Q, P, np_oracle.Obs('codes', S, n_sites, leaf_nodes=leaves, codes=codes), pi)
    groups, n_chains, burn, n_sweeps = 16, 256, 60, 100
    dwell = np.zeros((groups, S))
    trans = np.zeros((groups, S, S))
    for gidx in range(groups):
        ch = RaoTehChains(sched, Q, obs, n_chains=n_chains, root_distn=pi, seed=2000 + gidx)
        ch.sweep(burn, stats=False)
        ch.sweep(n_sweeps)
        ch.check()
        dwell[gidx] = ch.dwell_sum.cpu().numpy() / (n_chains * n_sweeps)
        trans[gidx] = ch.trans_sum.cpu().numpy() / (n_chains * n_sweeps)
    np.testing.assert_allclose(dwell.sum(axis=1), length.sum() * n_sites, rtol=1e-5)
    n_total = groups * n_chains * n_sweeps
    _assert_z(dwell, o['dwell'], n_total)
    _assert_z(trans.reshape(groups, -1), o['trans'].reshape(-1), n_total)
    # structure of a few histories
    ns, edges = ch.trajectory(3)
    for c, (times, states) in edges.items():
        assert states[0] == ns[parent[c]] and states[-1] == ns[c]
        assert np.all(states[1:] != states[:-1])
        for a, b in zip(states[:-1], states[1:]):
            assert Q[a, b] > 0


def _assert_z(got, want, n_total, zmax=5.0):
    """|mean - want| < zmax * se; se from the spread over groups, floored by the Poisson
    error sqrt(want / n_total) so that rare events (a handful of counts in the whole run,
    possibly none) are judged fairly."""
    groups = got.shape[0]
    mean = got.mean(axis=0)
    se = got.std(axis=0, ddof=1) / np.sqrt(groups)
    for m, s, w in zip(mean, se, want):
        if w == 0:
            assert m == 0, (m, w)
        else:
            s = max(s, np.sqrt(abs(w) / n_total))
            assert abs(m - w) < zmax * s + 1e-9, (m, w, s)


# ---------------------------------------------------------------------------------------
# K8: blocked Gibbs sampler of the tolerance process
# ---------------------------------------------------------------------------------------
def _toy_chains(n_chains, seed, disease, rate_on=0.7, rate_off=1.3, cap_p=64, cap_t=64):
    from raoteh_b200 import engine
    from raoteh_b200.tmjp import ToleranceChains
    Q, pi, part = toy_model()
    sched = toy_tree()
    node_to_state = {3: 4, 4: 5, 5: 1}
    leaves = np.array([3, 4, 5])
    codes = np.array([[4], [5], [1]], dtype=np.uint8)
    obs = engine.Observations.from_leaf_codes(sched, codes, leaf_nodes=leaves)
    tol_obs = tol_nodes = None
    if disease is not None:
        tol_nodes = sorted(set(n for d in disease for n in d))
        tol_obs = np.full((len(tol_nodes), 3, 1), 3, dtype=np.uint8)
        for c, d in enumerate(disease):
            for n, allowed in d.items():
                tol_obs[tol_nodes.index(n), c, 0] = sum(1 << s for s in allowed)
    ch = ToleranceChains(sched, Q, pi, dict(enumerate(part)), rate_on, rate_off, obs,
                         n_chains=n_chains, tol_obs=tol_obs, tol_obs_nodes=tol_nodes,
                         cap_p=cap_p, cap_t=cap_t, seed=seed)
    return ch, sched, Q, pi, part, node_to_state


DISEASE = [{5: {1}, 3: {0}}, {4: {0}}, {5: {0, 1}}]


@pytest.mark.parametrize('disease', [None, DISEASE])
def test_tolerance_histories_respect_structure(disease):
    ch, sched, Q, pi, part, node_to_state = _toy_chains(64, 5, disease)
    ch.initialize()
    ch.sweep(20)
    for t in range(0, ch.n_traj, 7):
        ns, edges = ch.primary_trajectory(t)
        for v, s in node_to_state.items():
            assert ns[v] == s
        tol = [ch.tolerance_trajectory(t, c) for c in range(3)]
        for c, (times, states) in edges.items():
            assert states[0] == ns[sched.parent[c]] and states[-1] == ns[c]
            assert np.all(states[1:] != states[:-1])
            assert np.all(np.diff(times) > 0) or len(times) < 2
            for a, b in zip(states[:-1], states[1:]):
                assert Q[a, b] > 0
            # compatibility: the class of the primary state is ON throughout every segment
            bounds = np.concatenate([[0.0], times, [sched.length[c]]])
            for i, s in enumerate(states):
                bits, tedges = tol[part[s]]
                tt, ts = tedges[c]
                tb = np.concatenate([[0.0], tt, [sched.length[c]]])
                for j, on in enumerate(ts):
                    overlap = min(bounds[i + 1], tb[j + 1]) - max(bounds[i], tb[j])
                    if overlap > 1e-7:
                        assert on == 1, (t, c, i, j)
        for cls in range(3):
            bits, tedges = tol[cls]
            for c, (tt, ts) in tedges.items():
                assert ts[0] == bits[sched.parent[c]] and ts[-1] == bits[c]
                assert len(ts) == len(tt) + 1
            if disease is not None:
                for v, allowed in disease[cls].items():
                    assert bits[v] in allowed


@pytest.mark.parametrize('disease', [None, DISEASE])
def test_tolerance_gibbs_matches_target_law(disease):
    """Means over 16 groups x 512 chains x 100 sweeps vs the closed form of the sampler's
    target law (oracle/np_tmjp.py), |z| < 5 per statistic."""
    ch0, sched, Q, pi, part, node_to_state = _toy_chains(1, 0, disease)
    want = np_tmjp.expected_sampler_statistics(
        sched.parent, sched.length, Q, part, 3, pi, 0.7, 1.3, node_to_state, disease)
    groups, n_chains, burn, n_sweeps = 16, 512, 80, 100
    dwell = np.zeros((groups, 6))
    trans = np.zeros((groups, 36))
    tol = np.zeros((groups, 12))
    for gidx in range(groups):
        ch = _toy_chains(n_chains, 4000 + gidx, disease)[0]
        ch.sweep(burn, stats=False)
        ch.sweep(n_sweeps)
        norm = n_chains * n_sweeps
        dwell[gidx] = ch.prim_dwell.cpu().numpy() / norm
        trans[gidx] = ch.prim_trans.cpu().numpy().ravel() / norm
        tol[gidx] = ch.tol_stats.cpu().numpy().ravel() / norm
    np.testing.assert_allclose(dwell.sum(axis=1), sched.length.sum(), rtol=1e-5)
    n_total = groups * n_chains * n_sweeps
    _assert_z(dwell, want['prim_dwell'], n_total)
    _assert_z(trans, want['prim_trans'].ravel(), n_total)
    _assert_z(tol, want['tol'].ravel(), n_total)


def test_fused_summary_equals_separate_summary():
    ch, sched, Q, pi, part, node_to_state = _toy_chains(256, 77, None)
    ch.initialize()
    ch.sweep(10, stats=False)
    ch.reset_statistics()
    from raoteh_b200 import tmjp
    ch._run(tmjp.MODE_SWEEP, n_sweeps=1, flags=tmjp.F_SUMMARY)     # summary fused into the sweep
    ch.sweeps_done += 1
    fused = ch.summary_out[:, :7].cpu().numpy().copy()
    fused_sum = ch.summary_sum.cpu().numpy().copy()
    ch.reset_statistics()
    sep = ch.tolerance_summary().cpu().numpy()
    np.testing.assert_allclose(fused, sep, rtol=1e-12)
    np.testing.assert_allclose(fused_sum[:7], sep.sum(axis=0), rtol=1e-10)
    assert fused_sum[7] == 256
    # identities of the summary (raoteh/sampler/_tmjp_dense.py:848-855)
    np.testing.assert_allclose(sep[:, 0] + sep[:, 1], 3.0, rtol=1e-12)
    np.testing.assert_allclose(sep[:, 2] + sep[:, 3], 3.0 * sched.length.sum(), rtol=1e-12)


def test_tolerance_sampler_is_counter_based():
    import torch
    full = _toy_chains(24, 9, DISEASE)[0]
    full.sweep(6)
    lo = _toy_chains(24, 9, DISEASE)[0]
    lo.n_traj = 10
    hi = _toy_chains(24, 9, DISEASE)[0]
    hi.traj0, hi.n_traj = 10, 14
    for p in (lo, hi):
        p.sweep(6)
    assert bool((torch.cat([lo.p_node[:10], hi.p_node[:14]]) == full.p_node).all())
    assert bool((torch.cat([lo.t_node[:10], hi.t_node[:14]]) == full.t_node).all())
    assert bool((torch.cat([lo.p_total[:10], hi.p_total[:14]]) == full.p_total).all())


# ---------------------------------------------------------------------------------------
# reference-shaped generators (raoteh_b200.sampler._sample_tmjp_dense / _sample_tmjp)
# ---------------------------------------------------------------------------------------
def _check_history(T, root, primary, tolerance, part_of, node_to_state, disease, total):
    import networkx as nx
    np.testing.assert_allclose(primary.size(weight='weight'), total, rtol=1e-5)
    assert set(T) <= set(primary)
    for v, s in node_to_state.items():
        for nb in primary[v]:
            assert primary[v][nb]['state'] == s
    for c, G in enumerate(tolerance):
        np.testing.assert_allclose(G.size(weight='weight'), total, rtol=1e-5)
        assert set(T) <= set(G)
        for a, b in G.edges():
            assert G[a][b]['state'] in (0, 1)
        if disease is not None:
            for v, allowed in disease[c].items():
                for nb in G[v]:
                    assert G[v][nb]['state'] in allowed
    # compatibility on every base edge: lay the primary and the class trajectory side by side
    def pieces(G, a, b):
        path = nx.shortest_path(G, a, b)
        out, t = [], 0.0
        for x, y in zip(path[:-1], path[1:]):
            out.append((t, t + G[x][y]['weight'], G[x][y]['state']))
            t += G[x][y]['weight']
        return out
    for a, b in nx.bfs_edges(T, root):
        for lo, hi, s in pieces(primary, a, b):
            for tlo, thi, on in pieces(tolerance[part_of[s]], a, b):
                if min(hi, thi) - max(lo, tlo) > 1e-6:
                    assert on == 1


def test_gen_histories_v1_generator():
    import networkx as nx
    from raoteh_b200.sampler import _tmjp_dense, _sample_tmjp_dense
    Q, pi, part = toy_model()
    ctm = _tmjp_dense.CompoundToleranceModel(Q, pi, dict(enumerate(int(p) for p in part)), 0.7, 1.3)
    T = nx.Graph()
    for a, b, w in ((10, 11, 0.5), (11, 12, 0.7), (12, 13, 0.4), (12, 14, 0.9), (11, 15, 0.6)):
        T.add_edge(a, b, weight=w)
    node_to_state = {13: 4, 14: 5, 15: 1}
    disease = [{15: {1}, 13: {0}}, {14: {0}}, {}]
    for dd in (None, disease):
        n = 0
        for primary, tolerance in _sample_tmjp_dense.gen_histories_v1(
                ctm, T, 10, node_to_state, disease_data=dd, nhistories=12, seed=3):
            n += 1
            assert len(tolerance) == 3
            _check_history(T, 10, primary, tolerance, part, node_to_state, dd, T.size(weight='weight'))
            ids = set(primary) - set(T)
            assert all(i > 15 for i in ids)
        assert n == 12


def test_gen_histories_sparse_generator():
    import networkx as nx
    from raoteh_b200.sampler import _tmjp, _sample_tmjp
    Q, pi, part = toy_model()
    names = ['a', 'b', 'c', 'd', 'e', 'f']
    Qs = nx.DiGraph()
    for i in range(6):
        for j in range(6):
            if i != j and Q[i, j] > 0:
                Qs.add_edge(names[i], names[j], weight=Q[i, j])
    ctm = _tmjp.CompoundToleranceModel(Qs, dict(zip(names, pi)), dict(zip(names, (int(p) for p in part))),
                                       0.7, 1.3)
    T = nx.Graph()
    for a, b, w in ((0, 1, 0.5), (1, 2, 0.7), (2, 3, 0.4), (2, 4, 0.9), (1, 5, 0.6)):
        T.add_edge(a, b, weight=w)
    node_to_state = {3: 'e', 4: 'f', 5: 'b'}
    part_of = dict(zip(names, (int(p) for p in part)))
    n = 0
    for primary, tolerance in _sample_tmjp.gen_histories(ctm, T, 0, node_to_state, nhistories=6, seed=8):
        n += 1
        _check_history(T, 0, primary, tolerance, part_of, node_to_state, None, T.size(weight='weight'))
        out = _tmjp.get_tolerance_summary(ctm, primary, 0)
        assert len(out) == 7 and abs(out[0] + out[1] - 3) < 1e-9
    assert n == 6


def test_importance_weights_estimate_the_likelihood_ratio():
    """raoteh/sampler/tests/test_sample_tmjp.py:186-246, :384-393: primary histories proposed by
    plain Rao-Teh under the proposal rate matrix, weight = exp(compound log-likelihood with the
    tolerance histories integrated out - proposal log-likelihood), everything on the device; the
    mean weight estimates P_compound(data) / P_proposal(data), computed here in closed form on
    the 48-state compound space.  |z| < 5."""
    import torch
    from raoteh_b200 import engine
    from raoteh_b200.raoteh import RaoTehChains
    from raoteh_b200.tmjp import ToleranceChains
    from raoteh_b200.sampler import _tmjp_dense
    Q, pi, part = toy_model()
    sched = toy_tree()
    rate_on, rate_off = 0.7, 1.3
    ctm = _tmjp_dense.CompoundToleranceModel(Q, pi, dict(enumerate(int(p) for p in part)), rate_on, rate_off)
    ctm.init_compound()
    Qp = _tmjp_dense.get_primary_proposal_rate_matrix(Q, ctm.primary_to_part, ctm.tolerance_distn)
    node_to_state = {3: 4, 4: 5, 5: 1}
    leaves = np.array([3, 4, 5])
    codes = np.array([[4], [5], [1]], dtype=np.uint8)
    # closed-form marginal likelihoods of the leaf data
    P = np_oracle.expm_edges(Qp, sched.length)
    ll_prop, _ = np_oracle.log_likelihood(sched.parent, P, np_oracle.Obs('codes', 6, 1, leaf_nodes=leaves, codes=codes), pi)
    nc = ctm.ncompound
    lik = np.ones((sched.n, 1, nc))
    for v, s in node_to_state.items():
        lik[v, 0] = [1.0 if p == s else 0.0 for p in ctm.compound_to_primary]
    Pc = np_oracle.expm_edges(ctm.Q_compound, sched.length)
    ll_comp, _ = np_oracle.log_likelihood(sched.parent, Pc, np_oracle.Obs('dense', nc, 1, lik=lik, has=np.ones(sched.n, dtype=bool)),
                                          ctm.compound_distn)
    want = float(np.exp(ll_comp[0] - ll_prop[0]))
    obs = engine.Observations.from_leaf_codes(sched, codes, leaf_nodes=leaves)
    groups, n_chains = 12, 4096
    means = []
    for g in range(groups):
        prop = RaoTehChains(sched, Qp, obs, n_chains=n_chains, root_distn=pi, seed=600 + g, cap=64)
        prop.sweep(60, stats=False)
        tol = ToleranceChains(sched, Q, pi, ctm.primary_to_part, rate_on, rate_off, obs, n_chains=n_chains)
        tol.attach_primary(prop)
        w = torch.exp(tol.tolerance_log_likelihood() - prop.trajectory_log_likelihood())
        means.append(float(w.mean()))
    m, se = np.mean(means), np.std(means, ddof=1) / np.sqrt(groups)
    assert abs(m - want) < 5 * se, (m, want, se)
    assert se < 0.05 * want


def test_device_metropolis_hastings_targets_the_compound_process():
    """Rao-Teh proposals under the approximate primary process + MH correction with the
    tolerance histories integrated out (raoteh/sampler/tests/test_sample_tmjp.py:248-276), all
    trajectories at once on the device: the primary dwell times and transition counts must
    match the exact posterior expectations of the 48-state compound process.  |z| < 5."""
    from raoteh_b200 import engine
    from raoteh_b200.mh import ToleranceMetropolisChains
    from raoteh_b200.sampler import _tmjp_dense
    Q, pi, part = toy_model()
    sched = toy_tree()
    rate_on, rate_off = 0.7, 1.3
    ctm = _tmjp_dense.CompoundToleranceModel(Q, pi, dict(enumerate(int(p) for p in part)), rate_on, rate_off)
    ctm.init_compound()
    node_to_state = {3: 4, 4: 5, 5: 1}
    leaves = np.array([3, 4, 5])
    codes = np.array([[4], [5], [1]], dtype=np.uint8)
    nc = ctm.ncompound
    lik = np.ones((sched.n, 1, nc))
    for v, s in node_to_state.items():
        lik[v, 0] = [1.0 if p == s else 0.0 for p in ctm.compound_to_primary]
    Pc = np_oracle.expm_edges(ctm.Q_compound, sched.length)
    r = np_oracle.expected_history_statistics(
        sched.parent, sched.length, ctm.Q_compound, Pc,
        np_oracle.Obs('dense', nc, 1, lik=lik, has=np.ones(sched.n, dtype=bool)), ctm.compound_distn)
    prim = np.array(ctm.compound_to_primary)
    want_dwell = np.array([r['dwell'][prim == p].sum() for p in range(6)])
    want_trans = np.zeros((6, 6))
    for i in range(nc):
        for j in range(nc):
            if prim[i] != prim[j]:
                want_trans[prim[i], prim[j]] += r['trans'][i, j]
    obs = engine.Observations.from_leaf_codes(sched, codes, leaf_nodes=leaves)
    groups, n_chains, burn, n_steps = 12, 1024, 40, 60
    dwell = np.zeros((groups, 6))
    trans = np.zeros((groups, 36))
    acc = []
    for g in range(groups):
        mh = ToleranceMetropolisChains(sched, Q, pi, ctm.primary_to_part, rate_on, rate_off, obs,
                                       n_chains=n_chains, seed=900 + g, cap=64)
        mh.step(burn, stats=False)
        mh.step(n_steps)
        dwell[g] = mh.dwell_sum.cpu().numpy() / (n_chains * n_steps)
        trans[g] = mh.trans_sum.cpu().numpy().ravel() / (n_chains * n_steps)
        acc.append(mh.n_accepted / mh.n_proposed)
    assert 0.2 < np.mean(acc) < 0.999
    np.testing.assert_allclose(dwell.sum(axis=1), sched.length.sum(), rtol=1e-5)
    n_total = groups * n_chains * n_steps
    _assert_z(dwell, want_dwell, n_total)
    _assert_z(trans, want_trans.ravel(), n_total)


def test_samplers_on_polytomies_chains_and_maximum_sizes():
    """Structure of sampled histories on a tree with a polytomy, a chain of degree-2 nodes and a
    zero-length branch, for the plain sampler with 64 states and the tolerance sampler with 64
    primary states in 32 classes (the maxima of the warp kernels)."""
    from raoteh_b200 import engine
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
    from raoteh_b200.tmjp import ToleranceChains
    parent = np.array([-1, 0, 1, 1, 1, 0, 5, 6, 6, 0], dtype=np.int32)
    rng = np.random.default_rng(12)
    length = rng.uniform(0.1, 0.5, size=len(parent))
    length[0] = 0.0
    length[3] = 0.0
    S = 64
    Q = rng.exponential(1.0, size=(S, S)) * (rng.random((S, S)) < 0.2)
    Q += np.roll(np.eye(S), 1, axis=1) + np.roll(np.eye(S), -1, axis=1)
    np.fill_diagonal(Q, 0)
    Q -= np.diag(Q.sum(axis=1))
    Q /= np.abs(np.diag(Q)).mean()
    pi = np.ones(S) / S
    sched = TreeSchedule(parent, length)
    leaves = sched.leaves
    codes = rng.integers(0, S, size=(len(leaves), 3)).astype(np.uint8)
    codes[0, 1] = 255
    obs = engine.Observations.from_leaf_codes(sched, codes, leaves)
    ch = RaoTehChains(sched, Q, obs, n_chains=5, root_distn=pi, seed=2, cap=400)
    ch.sweep(10)
    for t in range(ch.n_traj):
        ns, edges = ch.trajectory(t)
        assert len(edges[3][0]) == 0 and ns[3] == ns[1]          # nothing happens on a zero-length branch
        for c, (times, states) in edges.items():
            assert states[0] == ns[parent[c]] and states[-1] == ns[c]
            assert np.all(states[1:] != states[:-1])
            for a, b in zip(states[:-1], states[1:]):
                assert Q[a, b] > 0
        for i, leaf in enumerate(leaves):
            if codes[i, t % 3] != 255:
                assert ns[leaf] == codes[i, t % 3]
    part = dict((s, s % 32) for s in range(S))
    tc = ToleranceChains(sched, Q, pi, part, 0.8, 1.1, obs, n_chains=4, cap_p=400, cap_t=64, seed=5)
    tc.sweep(8, summary=True)
    for t in range(0, tc.n_traj, 5):
        ns, edges = tc.primary_trajectory(t)
        for c, (times, states) in edges.items():
            bounds = np.concatenate([[0.0], times, [length[c]]])
            for i, s in enumerate(states):
                bits, tedges = tc.tolerance_trajectory(t, part[int(s)])
                tt, ts = tedges[c]
                tb = np.concatenate([[0.0], tt, [length[c]]])
                for j, on in enumerate(ts):
                    if min(bounds[i + 1], tb[j + 1]) - max(bounds[i], tb[j]) > 1e-7:
                        assert on == 1
    out = tc.summary_out.cpu().numpy()
    np.testing.assert_allclose(out[:, 0] + out[:, 1], 32.0, rtol=1e-12)
    np.testing.assert_allclose(out[:, 2] + out[:, 3], 32.0 * length.sum(), rtol=1e-12)
    assert np.isfinite(out).all()


def test_codon_tolerance_gibbs_matches_cpu_port():
    """61 codons x 20 amino-acid classes (two states per lane, 20 of 32 class lanes): the closed
    form is out of reach (61 * 2^20 compound states), so the GPU sampler is compared with the CPU
    restatement of the same sweep (oracle/np_tmjp.py, itself checked against the closed form on
    the toy model).  Aggregate statistics, |z| < 5 with batch-means standard errors."""
    from raoteh_b200 import engine, synth
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.tmjp import ToleranceChains
    rng = np.random.default_rng(77)
    parent, length, leaves = synth.random_binary_tree(5, 0.15, rng)
    Q, pi, residues = synth.mg94(omega=1.0)
    aas = sorted(set(residues))
    part = np.array([aas.index(r) for r in residues])
    rate_on, rate_off = 0.21925, 0.78075
    Qp = synth.tolerance_proposal(Q, part, rate_on)
    codes = synth.simulate_leaf_codes(parent, length, leaves, Qp, pi, 1, rng, 0.0)
    sched = TreeSchedule(parent, length)
    node_to_state = dict((int(v), int(codes[i, 0])) for i, v in enumerate(leaves))
    human = int(leaves[0])
    dd = [dict() for _ in range(20)]
    tol_obs = np.full((1, 20, 1), 2, dtype=np.uint8)
    for c in (3, 7, 11):
        if c != part[codes[0, 0]]:
            dd[c][human] = {0}
            tol_obs[0, c, 0] = 1
    for c in range(20):
        dd[c].setdefault(human, {1})

    def summarise(dwell, trans, tol):
        same = part[:, None] == part[None, :]
        return np.array([trans[same].sum(), trans[~same].sum(), tol[:, 0].sum(), tol[:, 1].sum(),
                         tol[:, 2].sum(), tol[:, 3].sum(), (dwell * (-np.diag(Q))).sum()])
    # CPU port: one long chain, batch means
    crng = np.random.default_rng(5)
    margs = (parent, length, Q, part, 20, pi, rate_on, rate_off, node_to_state, dd)
    prim, tols = np_tmjp.gibbs_init(*margs, crng)
    burn, nb, per = 100, 20, 60
    rows = []
    for i in range(burn + nb * per):
        prim, tols = np_tmjp.gibbs_sweep(*margs, prim, tols, crng)
        if i >= burn:
            rows.append(summarise(*np_tmjp.sampled_statistics(parent, length, prim, tols, 61, 20)))
    rows = np.array(rows).reshape(nb, per, -1).mean(axis=1)
    cpu_mean, cpu_se = rows.mean(axis=0), rows.std(axis=0, ddof=1) / np.sqrt(nb)
    # GPU: independent groups of chains
    obs = engine.Observations.from_leaf_codes(sched, codes, leaves)
    groups, n_chains, n_sweeps = 12, 256, 60
    g = []
    for k in range(groups):
        ch = ToleranceChains(sched, Q, pi, dict(enumerate(int(p) for p in part)), rate_on, rate_off, obs,
                             n_chains=n_chains, tol_obs=tol_obs, tol_obs_nodes=[human], cap_p=96,
                             cap_t=48, seed=300 + k)
        ch.sweep(80, stats=False)
        ch.sweep(n_sweeps)
        norm = n_chains * n_sweeps
        g.append(summarise(ch.prim_dwell.cpu().numpy() / norm, ch.prim_trans.cpu().numpy() / norm,
                           ch.tol_stats.cpu().numpy() / norm))
    g = np.array(g)
    gpu_mean, gpu_se = g.mean(axis=0), g.std(axis=0, ddof=1) / np.sqrt(groups)
    for a, sa, b, sb in zip(gpu_mean, gpu_se, cpu_mean, cpu_se):
        assert abs(a - b) < 5 * np.hypot(sa, sb) + 1e-9, (a, sa, b, sb)


@pytest.mark.parametrize('level', ['L1', 'L2'])
def test_code2x3_blinking_model_golden_values_by_sampling(level):
    """The published numbers of the blinking model of examples/code2x3
    (full-description.tex:305-370; inputs run.py:520-614): expected synonymous / non-synonymous
    primary transitions and tolerance gains / losses summed over the five branches, with
    alignment data only (L1) and with disease data at the root as well (L2).  Reproduced here by
    SAMPLING: Rao-Teh proposals under the approximate primary process, Metropolis-Hastings
    correction with the tolerance histories integrated out, Rao-Blackwellised tolerance summary
    of every current history -- all on the device.  |z| < 5."""
    from raoteh_b200 import engine
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.mh import ToleranceMetropolisChains
    Q, pi, part = toy_model()            # the 6-state code of run.py:524-540, classes {0,1},{2,3},{4,5}
    parent = np.array([-1, 0, 1, 2, 2, 1], dtype=np.int32)       # run.py:557-559
    sched = TreeSchedule(parent, np.array([0.0, 0.5, 0.5, 0.5, 0.5, 0.5]))
    obs_nodes = np.array([0, 3, 4, 5])                            # run.py:578-585
    codes = np.array([[0], [4], [5], [1]], dtype=np.uint8)
    obs = engine.Observations.from_leaf_codes(sched, codes, leaf_nodes=obs_nodes)
    golden = dict(     # per-branch values of the tex file, summed over the branches
        L1=dict(syn=4 * 0.530243839876 + 0.214285714286, nonsyn=4 * 0.214361006829 + 1.63974144573,
                gain=0.679310052688 + 0.540928083993 + 3 * 0.316402335024,
                loss=0.316402335024 + 0.540928083993 + 3 * 0.679310052688),
        L2=dict(syn=4 * 0.530243839876 + 0.214285714286,
                nonsyn=0.115905720866 + 1.71225589525 + 2 * 0.22303974637 + 0.128868804275,
                gain=0.880141827661 + 0.563730372337 + 2 * 0.307552791569 + 0.315487814895,
                loss=0.30799029452 + 0.524525848196 + 2 * 0.685431320483 + 0.681655127317))[level]
    tol_obs = tol_nodes = None
    if level == 'L2':                                             # run.py:603-611: root (0,0):{1}, (0,1):{0}, (0,2):{1}
        tol_nodes = [0]
        tol_obs = np.array([[[2], [1], [2]]], dtype=np.uint8)
    same = part[:, None] == part[None, :]
    groups, n_chains, burn, n_steps = 12, 2048, 40, 50
    rows = []
    for g in range(groups):
        mh = ToleranceMetropolisChains(sched, Q, pi, dict(enumerate(int(p) for p in part)), 1.0, 1.0, obs,
                                       n_chains=n_chains, seed=1200 + g, cap=64, tol_obs=tol_obs,
                                       tol_obs_nodes=tol_nodes)
        mh.step(burn, stats=False)
        mh.step(n_steps)
        norm = n_chains * n_steps
        tr = mh.trans_sum.cpu().numpy() / norm
        sm = mh.summary_acc.cpu().numpy() / norm
        rows.append([tr[same].sum(), tr[~same].sum(), sm[5], sm[6]])
    rows = np.array(rows)
    m, se = rows.mean(axis=0), rows.std(axis=0, ddof=1) / np.sqrt(groups)
    want = np.array([golden['syn'], golden['nonsyn'], golden['gain'], golden['loss']])
    for a, s, w in zip(m, se, want):
        assert abs(a - w) < 5 * s, (level, m, se, want)
    assert (se < 0.02 * want).all()


def test_scalar_tolerance_process_log_likelihood_matches_reference_fixture():
    """_tmjp.get_tolerance_process_log_likelihood (raoteh/sampler/_tmjp.py:406-490) and its dense
    twin (_tmjp_dense.py:407-505), the scalar mirrors: against the values the reference itself
    returned for the fixture trajectories (relative 1e-9: the 3 x 3 pieces go through the generic
    fp64 path, rt_expm_batched + rt_prune_loglik with one rate matrix per edge)."""
    import networkx as nx
    from raoteh_b200.sampler import _tmjp, _tmjp_dense
    g = load_golden('tolerance_summary.json')
    Q = np.array(g['Q_primary'])
    part = dict((i, g['primary_to_part'][str(i)]) for i in range(6))
    distn = np.ones(6) / 6
    Qs = nx.DiGraph()
    for a in range(6):
        for b in range(6):
            if a != b and Q[a, b] > 0:
                Qs.add_edge(a, b, weight=Q[a, b])
    n = 0
    for case in g['cases']:
        if case['disease'] is not None or 'tol_ll' not in case:
            continue
        T = nx.Graph()
        for a, b, w, s in case['edges']:
            T.add_edge(a, b, weight=w, state=s)
        got = _tmjp_dense.get_tolerance_process_log_likelihood(
            Q, part, T, case['rate_off'], case['rate_on'], distn, case['root'])
        np.testing.assert_allclose(got, case['tol_ll'], rtol=1e-9)
        ctm = _tmjp.CompoundToleranceModel(Qs, dict(enumerate(distn)), part, case['rate_on'], case['rate_off'])
        got = _tmjp.get_tolerance_process_log_likelihood(ctm, T, case['root'])
        np.testing.assert_allclose(got, case['tol_ll'], rtol=1e-9)
        n += 1
        if n >= 6:
            break
    assert n >= 3
