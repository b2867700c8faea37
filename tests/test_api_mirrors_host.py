"""Host-side API mirrors that need no GPU: the stand-alone Poisson step, the forward simulators
and the differential-entropy helpers (raoteh/sampler/_sample_mjp.py:19, _sampler.py:67,
_mjp.py:255, _tmjp.py:217 / _tmjp_dense.py:508).  Where /root/reference is present the helpers
are compared with the reference's own functions on the same inputs."""
import networkx as nx
import numpy as np
import pytest


def _trajectory():
    T = nx.Graph()
    T.add_edge(0, 1, weight=0.5, state=0)
    T.add_edge(1, 2, weight=1.5, state=1)
    T.add_edge(1, 3, weight=0.25, state=2)
    T.add_edge(3, 7, weight=2.0, state=0)
    return T


def test_resample_poisson_structure_and_rate():
    from raoteh_b200.sampler import _sample_mjp
    T = _trajectory()
    rates = {0: 3.0, 1: 0.5, 2: 0.0}
    np.random.seed(5)
    counts = []
    for _ in range(400):
        T_out = _sample_mjp.resample_poisson(T, rates, root=0)
        assert set(T) <= set(T_out) and min(set(T_out) - set(T), default=8) > max(T)
        np.testing.assert_allclose(T_out.size(weight='weight'), T.size(weight='weight'), rtol=1e-12)
        assert all('state' not in d for _, _, d in T_out.edges(data=True))
        assert nx.is_tree(T_out)
        # original nodes keep their distances from the root
        dist = nx.single_source_dijkstra_path_length(T_out, 0)
        np.testing.assert_allclose([dist[1], dist[2], dist[3], dist[7]], [0.5, 2.0, 0.75, 2.75], rtol=1e-12)
        counts.append(T_out.number_of_nodes() - T.number_of_nodes())
    want = 3.0 * 0.5 + 0.5 * 1.5 + 0.0 + 3.0 * 2.0          # sum over segments of rate x length
    assert abs(np.mean(counts) - want) < 5 * np.sqrt(want / 400)
    # the dense twin takes the rates as an array
    T_out = _sample_mjp.resample_poisson_dense(T, np.array([3.0, 0.5, 0.0]), root=0)
    np.testing.assert_allclose(T_out.size(weight='weight'), 4.25, rtol=1e-12)


def test_gen_forward_samples():
    from raoteh_b200.sampler import _sampler
    T = nx.Graph()
    T.add_edge(0, 1, weight=1.0)
    T.add_edge(0, 2, weight=2.0)
    T.add_edge(2, 3, weight=0.5)
    Q = nx.DiGraph()
    Q.add_edge('a', 'b', weight=1.0)
    Q.add_edge('b', 'a', weight=2.0)
    np.random.seed(1)
    out = list(_sampler.gen_forward_samples(T, Q, 0, {'a': 0.5, 'b': 0.5}, nsamples=30))
    assert len(out) == 30
    njumps = 0
    for H in out:
        np.testing.assert_allclose(H.size(weight='weight'), 3.5, rtol=1e-12)
        assert all(d['state'] in ('a', 'b') and d['weight'] >= 0 for _, _, d in H.edges(data=True))
        njumps += H.number_of_nodes() - 4
    assert njumps > 20          # mean jump rate ~ 4/3 over tree length 3.5


def test_mjp_differential_entropy_helper():
    from scipy import special
    from raoteh_b200.sampler import _mjp
    Q = nx.DiGraph()
    Q.add_edge(0, 1, weight=0.5)
    Q.add_edge(1, 0, weight=2.0)
    Q.add_edge(1, 2, weight=0.25)
    prior = {0: 0.5, 1: 0.25, 2: 0.25}
    post_root = {0: 0.9, 1: 0.1}
    dwell = {0: 1.2, 1: 0.3, 2: 0.7}
    trans = nx.DiGraph()
    trans.add_edge(0, 1, weight=0.6)
    trans.add_edge(1, 2, weight=0.1)
    trans.add_edge(2, 0, weight=0.3)      # not a transition of Q: ignored
    init, dw, tr = _mjp.differential_entropy_helper(Q, prior, post_root, dwell, trans)
    np.testing.assert_allclose(init, -(0.9 * np.log(0.5) + 0.1 * np.log(0.25)))
    np.testing.assert_allclose(dw, 1.2 * 0.5 + 0.3 * 2.25)
    np.testing.assert_allclose(tr, -(special.xlogy(0.6, 0.5) + special.xlogy(0.1, 0.25)))


@pytest.mark.reference
def test_entropy_helpers_match_reference():
    """the three differential_entropy_helper mirrors against the reference's own functions"""
    from oracle import ref_shim
    ref_shim.load_reference()
    from raoteh.sampler import _mjp as r_mjp, _tmjp as r_tmjp, _tmjp_dense as r_tmjp_dense
    from raoteh_b200.sampler import _mjp, _tmjp, _tmjp_dense
    rng = np.random.default_rng(3)
    # plain MJP
    Q = nx.DiGraph()
    for a in range(4):
        for b in range(4):
            if a != b and rng.random() < 0.8:
                Q.add_edge(a, b, weight=float(rng.exponential()))
    prior = dict(enumerate(rng.dirichlet(np.ones(4))))
    post = dict(enumerate(rng.dirichlet(np.ones(4))))
    dwell = dict(enumerate(rng.exponential(size=4)))
    trans = nx.DiGraph()
    for a, b in Q.edges():
        trans.add_edge(a, b, weight=float(rng.exponential()))
    np.testing.assert_allclose(_mjp.differential_entropy_helper(Q, prior, post, dwell, trans),
                               r_mjp.differential_entropy_helper(Q, prior, post, dwell, trans), rtol=1e-13)
    # compound tolerance model: 4 primary states in 2 classes
    Qp = rng.exponential(size=(4, 4))
    np.fill_diagonal(Qp, 0)
    Qp -= np.diag(Qp.sum(axis=1))
    distn = rng.dirichlet(np.ones(4))
    part = {0: 0, 1: 0, 2: 1, 3: 1}
    mine = _tmjp_dense.CompoundToleranceModel(Qp, distn, part, 0.7, 1.3)
    mine.init_compound()
    theirs = r_tmjp_dense.CompoundToleranceModel(Qp, distn, part, 0.7, 1.3)
    theirs.init_compound()
    np.testing.assert_allclose(mine.Q_compound, theirs.Q_compound, rtol=1e-13)
    n = mine.ncompound
    ok = np.asarray(mine.compound_distn) > 0
    post_root = np.where(ok, rng.random(n), 0.0)
    post_root /= post_root.sum()
    dwell = np.where(ok, rng.exponential(size=n), 0.0)
    trans = np.where(np.asarray(mine.Q_compound) > 0, rng.exponential(size=(n, n)), 0.0)
    a = _tmjp_dense.differential_entropy_helper(mine, post_root, dwell, trans)
    b = r_tmjp_dense.differential_entropy_helper(theirs, post_root, dwell, trans)
    for k in ('init_prim', 'init_tol', 'dwell_prim', 'dwell_tol', 'trans_prim', 'trans_tol'):
        np.testing.assert_allclose(getattr(a, k), getattr(b, k), rtol=1e-12, err_msg=k)
    # sparse twin
    Qs = nx.DiGraph()
    for i in range(4):
        for j in range(4):
            if i != j:
                Qs.add_edge(i, j, weight=float(Qp[i, j]))
    s_mine = _tmjp.CompoundToleranceModel(Qs, dict(enumerate(distn)), part, 0.7, 1.3)
    s_theirs = r_tmjp.CompoundToleranceModel(Qs, dict(enumerate(distn)), part, 0.7, 1.3)
    s_theirs.init_compound()
    pr = dict((i, float(v)) for i, v in enumerate(post_root) if v)
    dw = dict((i, float(v)) for i, v in enumerate(dwell) if v)
    tg = nx.DiGraph()
    for i, j in zip(*np.nonzero(trans)):
        tg.add_edge(int(i), int(j), weight=float(trans[i, j]))
    a = _tmjp.differential_entropy_helper(s_mine, pr, dw, tg)
    b = r_tmjp.differential_entropy_helper(s_theirs, pr, dw, tg)
    for k in ('init_prim', 'init_tol', 'dwell_prim', 'dwell_tol', 'trans_prim', 'trans_tol'):
        np.testing.assert_allclose(getattr(a, k), getattr(b, k), rtol=1e-12, err_msg=k)
