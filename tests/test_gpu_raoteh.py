"""Rao-Teh sweeps (K6) on the GPU: structural invariants of every sampled history
(the assertions of raoteh/sampler/tests/test_sampler.py:398-438) and distributional
agreement with the closed-form posterior expectations
(_mjp.get_expected_history_statistics), which the reference only prints
(tests/test_sampler.py:251-393).  z-score bound stated per test."""
import numpy as np
import pytest

from oracle import np_oracle

pytestmark = pytest.mark.gpu


def _setup(S, n_leaves, n_sites, seed, missing=0.0):
    from raoteh_b200 import synth, engine
    from raoteh_b200.lowering import TreeSchedule
    rng = np.random.default_rng(seed)
    parent, length, leaves = synth.random_binary_tree(n_leaves, 0.25, rng)
    if S == 4:
        Q, pi = synth.hky85()
    else:
        Q = rng.exponential(1.0, size=(S, S))
        np.fill_diagonal(Q, 0)
        Q -= np.diag(Q.sum(axis=1))
        Q /= np.abs(np.diag(Q)).mean()
        pi = rng.dirichlet(np.ones(S) * 3)
    codes = synth.simulate_leaf_codes(parent, length, leaves, Q, pi, n_sites, rng, missing)
    sched = TreeSchedule(parent, length)
    obs = engine.Observations.from_leaf_codes(sched, codes, leaves)
    return parent, length, leaves, Q, pi, codes, sched, obs


def test_histories_respect_structure():
    from raoteh_b200.raoteh import RaoTehChains
    parent, length, leaves, Q, pi, codes, sched, obs = _setup(4, 12, 5, 11, missing=0.1)
    ch = RaoTehChains(sched, Q, obs, n_chains=7, root_distn=pi, seed=3)
    ch.initialize()
    ch.sweep(25)
    ch.check()
    for t in range(ch.n_traj):
        site = t % 5
        ns, edges = ch.trajectory(t)
        assert len(edges) == sched.n - 1
        for c, (times, states) in edges.items():
            assert states[0] == ns[parent[c]]            # consistent state around original nodes
            assert states[-1] == ns[c]
            assert len(states) == len(times) + 1
            assert np.all(np.diff(times) > 0) or len(times) < 2
            assert np.all(times > 0) and np.all(times < length[c])
            assert np.all(states[1:] != states[:-1])      # self-transitions were dropped
        for i, leaf in enumerate(leaves):                 # observed leaf states are respected
            if codes[i, site] != 255:
                assert ns[leaf] == codes[i, site]


def test_philox_is_counter_based():
    """Same (seed, global trajectory index, sweep) -> same history, however the
    trajectories are split over launches / ranks."""
    import torch
    from raoteh_b200.raoteh import RaoTehChains
    parent, length, leaves, Q, pi, codes, sched, obs = _setup(4, 10, 6, 5)
    full = RaoTehChains(sched, Q, obs, n_chains=4, root_distn=pi, seed=9)
    full.sweep(5)
    full.sweep(3)
    lo = RaoTehChains(sched, Q, obs, n_chains=4, root_distn=pi, seed=9, traj0=0, n_traj=10)
    hi = RaoTehChains(sched, Q, obs, n_chains=4, root_distn=pi, seed=9, traj0=10, n_traj=14)
    for part in (lo, hi):
        part.sweep(8)
    ns = torch.cat([lo.node_state, hi.node_state], dim=1)
    assert bool((ns == full.node_state).all())
    tot = torch.cat([lo.ev_total, hi.ev_total])
    assert bool((tot == full.ev_total).all())
    np.testing.assert_allclose((lo.dwell_sum + hi.dwell_sum).cpu().numpy(),
                               full.dwell_sum.cpu().numpy(), rtol=1e-12)


@pytest.mark.parametrize('S,n_leaves', [(4, 8), (3, 5), (6, 6)])
def test_sweep_statistics_match_closed_form(S, n_leaves):
    """Mean dwell times / transition counts over many chains and sweeps vs the
    closed-form posterior expectations; |z| < 5 per statistic with the standard
    error estimated from 16 independent groups of chains."""
    from raoteh_b200.raoteh import RaoTehChains
    n_sites = 3
    parent, length, leaves, Q, pi, codes, sched, obs = _setup(S, n_leaves, n_sites, 100 + S)
    P = np_oracle.expm_edges(Q, length)
    o = np_oracle.expected_history_statistics(
        parent, length, Q, P, np_oracle.Obs('codes', S, n_sites, leaf_nodes=leaves, codes=codes), pi)
    groups, n_chains, burn, n_sweeps = 16, 512, 60, 120
    dwell = np.zeros((groups, S))
    trans = np.zeros((groups, S, S))
    for gidx in range(groups):
        ch = RaoTehChains(sched, Q, obs, n_chains=n_chains, root_distn=pi, seed=1000 + gidx)
        ch.sweep(burn, stats=False)
        ch.sweep(n_sweeps)
        ch.check()
        dwell[gidx] = ch.dwell_sum.cpu().numpy() / (n_chains * n_sweeps)
        trans[gidx] = ch.trans_sum.cpu().numpy() / (n_chains * n_sweeps)
    # total dwell per sweep is exactly the tree length (float32 event times)
    np.testing.assert_allclose(dwell.sum(axis=1), length.sum() * n_sites, rtol=1e-5)
    for got, want in ((dwell, o['dwell']), (trans.reshape(groups, -1), o['trans'].reshape(-1))):
        mean = got.mean(axis=0)
        se = got.std(axis=0, ddof=1) / np.sqrt(groups)
        for m, s, w in zip(mean, se, want):
            if w == 0:
                assert m == 0
            else:
                assert abs(m - w) < 5 * s + 1e-9, (m, w, s)


def test_jukes_cantor_path_dwell():
    """raoteh/sampler/tests/test_sampler.py:86-122: one branch with known end states."""
    from raoteh_b200 import engine
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
    n, t = 4, 0.5
    Q = (np.ones((n, n)) - np.eye(n)) / (n - 1)
    Q -= np.diag(Q.sum(axis=1))
    sched = TreeSchedule(np.array([-1, 0], dtype=np.int32), np.array([0.0, t]))
    a, b = 0, 2
    mask = np.array([[1 << a], [1 << b]], dtype=np.uint64)
    obs = engine.Observations.from_masks(sched, mask)
    ch = RaoTehChains(sched, Q, obs, n_chains=20000, seed=4)
    ch.sweep(30, stats=False)
    ch.sweep(50)
    ch.check()
    p = np.exp(-(n * t) / (n - 1))
    pm1 = np.expm1(-(n * t) / (n - 1))
    pab = (1 - p) / n

    def interaction(c):   # _conditional_expectation.py:35-46 with d = c
        if a != c and c != b:
            x = t * p + pm1 * 2 * (n - 1) / n
        else:
            x = -(n - 1) * t * p - pm1 * (n - 2) * (n - 1) / n
        return (t + x) / (n * n)
    want = np.array([interaction(c) / pab for c in range(n)])
    got = ch.dwell_sum.cpu().numpy() / (20000 * 50)
    np.testing.assert_allclose(got, want, rtol=0.02)


def test_infeasible_data_is_reported():
    """raoteh/sampler/tests/test_sample_mcx.py:79-99 analogue: no feasible history."""
    from raoteh_b200 import engine
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
    Q = np.array([[-1.0, 1.0, 0.0], [0.0, -1.0, 1.0], [0.0, 0.0, 0.0]])   # 0 -> 1 -> 2, absorbing
    sched = TreeSchedule(np.array([-1, 0], dtype=np.int32), np.array([0.0, 1.0]))
    mask = np.array([[1 << 2], [1 << 0]], dtype=np.uint64)                # root in 2, leaf in 0
    obs = engine.Observations.from_masks(sched, mask)
    ch = RaoTehChains(sched, Q, obs, n_chains=3, seed=1)
    with pytest.raises(RuntimeError):
        ch.initialize()


def test_mh_histories_with_exact_target_always_accepts_and_biased_target_rejects():
    """raoteh/sampler/_sampler.py:393-551.  With target == proposal the MH ratio is 1 and
    every proposal is accepted; with a target that penalises transitions some proposals are
    rejected, a rejected step re-yields the previous history object, and the sample mean of
    the number of transitions drops."""
    import functools
    import random
    import networkx as nx
    from raoteh_b200.sampler import _sampler, _mjp
    random.seed(5)
    Q = nx.DiGraph()
    for a, b, w in ((0, 1, 1.0), (1, 0, 0.5), (1, 2, 0.7), (2, 1, 0.9), (2, 0, 0.3), (0, 2, 0.4)):
        Q.add_edge(a, b, weight=w)
    T = nx.Graph()
    for a, b, w in ((0, 1, 0.8), (1, 2, 0.6), (1, 3, 1.1), (0, 4, 0.9)):
        T.add_edge(a, b, weight=w)
    allowed = {0: {0, 1, 2}, 1: {0, 1, 2}, 2: {0}, 3: {2}, 4: {1}}
    distn = {0: 0.3, 1: 0.3, 2: 0.4}
    exact = functools.partial(_mjp.get_trajectory_log_likelihood, root=0, prior_root_distn=distn,
                              Q_default=Q)
    flags = [f for _, f in _sampler.gen_mh_histories(T, Q, allowed, exact, 0, root_distn=distn,
                                                     nhistories=40, seed=1)]
    assert all(flags)

    def ntrans(T_aug):
        return sum(1 for v in T_aug if v not in T)

    def penalised(T_aug):
        return exact(T_aug) - 1.5 * ntrans(T_aug)
    prev, counts, rejected = None, [], 0
    for T_aug, flag in _sampler.gen_mh_histories(T, Q, allowed, penalised, 0, root_distn=distn,
                                                 nhistories=300, seed=2):
        assert abs(T_aug.size(weight='weight') - T.size(weight='weight')) < 1e-5
        if not flag:
            rejected += 1
            assert T_aug is prev
        prev = T_aug
        counts.append(ntrans(T_aug))
    assert 10 < rejected < 290
    plain = [ntrans(h) for h in _sampler.gen_restricted_histories(T, Q, allowed, 0, root_distn=distn,
                                                                   nhistories=300, seed=3)]
    assert np.mean(counts[50:]) < np.mean(plain[50:])


def test_birth_death_indel_bridge():
    """raoteh/sampler/tests/test_sample_mjp.py:29-112 (a print-only, disabled test in the
    reference): sequence-length birth-death process with 50 states on one branch, end states
    3 and 7.  Every sampled history has an even number of excess events, and the mean number
    of transitions matches the closed-form posterior expectation (|z| < 5)."""
    from raoteh_b200 import engine
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
    S, t = 50, 0.5      # (the reference uses t = 2; here omega * t stays below the 255 events per branch)
    Q = np.zeros((S, S))                      # state index i = sequence length i + 1
    for i in range(S):
        if i > 0:
            Q[i, i - 1] = i * 1.5
        if i < S - 1:
            Q[i, i + 1] = (i + 1) * 1.0
    Q -= np.diag(Q.sum(axis=1))
    a, b = 3 - 1, 7 - 1
    parent = np.array([-1, 0], dtype=np.int32)
    length = np.array([0.0, t])
    sched = TreeSchedule(parent, length)
    mask = np.array([[1 << a], [1 << b]], dtype=np.uint64)
    obs = engine.Observations.from_masks(sched, mask)
    P = np_oracle.expm_edges(Q, length)
    o = np_oracle.expected_history_statistics(parent, length, Q, P, np_oracle.Obs('mask', S, 1, mask=mask), None)
    want = o['trans'].sum()
    groups, n_chains, n_sweeps = 12, 512, 60
    means = []
    for g in range(groups):
        ch = RaoTehChains(sched, Q, obs, n_chains=n_chains, seed=70 + g, cap=400)
        ch.sweep(60, stats=False)
        ch.sweep(n_sweeps)
        means.append(float(ch.trans_sum.sum()) / (n_chains * n_sweeps))
        tot = ch.ev_total.cpu().numpy()
        assert np.all((tot - (b - a)) % 2 == 0) and np.all(tot >= b - a)
    m, se = np.mean(means), np.std(means, ddof=1) / np.sqrt(groups)
    assert abs(m - want) < 5 * se, (m, want, se)


def test_event_capacity_overflow_is_loud():
    """More candidate events than a branch can hold (255) must raise, not return stale histories."""
    from raoteh_b200 import engine, _native
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
    Q = np.array([[-200.0, 200.0], [200.0, -200.0]])
    sched = TreeSchedule(np.array([-1, 0], dtype=np.int32), np.array([0.0, 2.0]))
    obs = engine.Observations.from_masks(sched, np.array([[1], [1]], dtype=np.uint64))
    ch = RaoTehChains(sched, Q, obs, n_chains=4, seed=1, cap=4000)
    with pytest.raises(_native.NativeError):
        ch.sweep(2)


def test_capacity_growth_keeps_histories():
    """grow(): larger pools, same histories; auto_grow continues after an overflow."""
    import torch
    from raoteh_b200.raoteh import RaoTehChains
    parent, length, leaves, Q, pi, codes, sched, obs = _setup(4, 10, 4, 21)
    a = RaoTehChains(sched, Q, obs, n_chains=8, root_distn=pi, seed=3, cap=96)
    b = RaoTehChains(sched, Q, obs, n_chains=8, root_distn=pi, seed=3, cap=96)
    a.sweep(5)
    b.sweep(5)
    b.grow(160)
    a.sweep(4)
    b.sweep(4)
    assert torch.equal(a.node_state, b.node_state) and torch.equal(a.ev_total, b.ev_total)
    for t in (0, 7, 31):
        ea, eb = a.trajectory(t)[1], b.trajectory(t)[1]
        for c in ea:
            np.testing.assert_array_equal(ea[c][0], eb[c][0])
            np.testing.assert_array_equal(ea[c][1], eb[c][1])
    tiny = RaoTehChains(sched, Q, obs, n_chains=8, root_distn=pi, seed=3, cap=12)
    tiny.initialize()
    tiny.sweep(3, auto_grow=True)
    assert tiny.cap > 12 and int((tiny.status != 0).sum()) == 0


def test_auto_grow_keeps_every_trajectory_on_the_same_sweep():
    """A trajectory that runs out of event capacity stops at its last completed sweep; with
    auto_grow the pools are doubled and the SAME call is repeated, the stopped trajectories resume
    at their own sweep index: afterwards every trajectory has done every sweep exactly once (the
    statistics hold n_traj x n_sweeps histories: total dwell = tree length x that count) and the
    histories are those of a run that had the larger capacity from the start."""
    import torch
    from raoteh_b200.raoteh import RaoTehChains
    parent, length, leaves, Q, pi, codes, sched, obs = _setup(4, 12, 4, 21)
    small = RaoTehChains(sched, Q, obs, n_chains=300, root_distn=pi, seed=8, cap=24)
    small.initialize()
    n_sweeps = 40
    small.sweep(n_sweeps, auto_grow=True)
    assert small.cap > 24                                  # the capacity was in fact exceeded
    assert bool((small.sweep_count == 1 + n_sweeps).all())
    np.testing.assert_allclose(float(small.dwell_sum.sum()), length.sum() * small.n_traj * n_sweeps, rtol=1e-5)
    big = RaoTehChains(sched, Q, obs, n_chains=300, root_distn=pi, seed=8, cap=small.cap)
    big.initialize()
    big.sweep(n_sweeps)
    assert bool((big.node_state == small.node_state).all())
    assert bool((big.ev_total == small.ev_total).all())
    assert torch.equal(big.trans_sum, small.trans_sum)


def test_long_event_chains_are_subdivided():
    """_sample_mcy.resample_edge_states on a path with more event nodes than the kernels' uint8
    per-branch counters hold (the reference takes any number, raoteh/sampler/_sample_mcy.py:86-187):
    the chain is cut with unrestricted pseudo nodes; load_events itself refuses > 255 loudly."""
    import networkx as nx
    from raoteh_b200 import engine
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
    from raoteh_b200.sampler import _sample_mcy
    n_ev = 700
    T = nx.Graph()
    for i in range(n_ev + 1):
        T.add_edge(i, i + 1, weight=1.0)
    root, leaf = 0, n_ev + 1
    P = nx.DiGraph()
    for a in range(3):
        for b in range(3):
            P.add_edge(a, b, weight=0.8 if a == b else 0.1)
    T_aug = _sample_mcy.resample_edge_states(T, root, P, set(range(1, n_ev + 1)),
                                             node_to_allowed_states={root: {0}, leaf: {2}}, seed=5)
    states = [T_aug[i][i + 1]['state'] for i in range(n_ev + 1)]
    assert states[0] == 0 and states[-1] == 2 and T_aug.number_of_edges() == n_ev + 1
    changes = sum(1 for a, b in zip(states[:-1], states[1:]) if a != b)
    assert 60 < changes < 220                # ~ Binomial(700, 0.2) conditioned on the end states
    sched = TreeSchedule(np.array([-1, 0], dtype=np.int32), np.array([0.0, 1.0]))
    obs = engine.Observations.from_masks(sched, np.array([[7], [7]], dtype=np.uint64))
    ch = RaoTehChains(sched, None, obs, n_chains=1, cap=400, chain_matrix=np.full((3, 3), 1 / 3.0))
    with pytest.raises(ValueError):
        ch.load_events({1: list(np.linspace(0.001, 0.999, 300))})
