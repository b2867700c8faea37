"""The reference-shaped API (raoteh_b200.sampler) against golden fixtures produced
by the unmodified reference, plus ports of the reference's own tests
(raoteh/sampler/tests/test_mjp.py, test_sampler.py:398-438)."""
from itertools import product

import networkx as nx
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_equal

from helpers import load_golden, case_tree, case_allowed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def S():
    from raoteh_b200 import sampler  # noqa: F401
    from raoteh_b200.sampler import (_util, _density, _mjp, _mjp_dense, _mcy, _mcy_dense, _mcz,
                                      _mc0, _mc0_dense, _sampler, _sample_mjp,
                                      _conditional_expectation)

    class NS(object):
        pass
    ns = NS()
    for m in (_util, _density, _mjp, _mjp_dense, _mcy, _mcy_dense, _mcz, _mc0, _mc0_dense,
              _sampler, _sample_mjp, _conditional_expectation):
        setattr(ns, m.__name__.split('.')[-1], m)
    return ns


def test_dense_api_matches_reference_fixtures(S):
    g = load_golden('mjp_random.json')
    for case in g['cases']:
        T, root = case_tree(case)
        allowed = case_allowed(case)
        n = case['nstates']
        Q = np.array(case['Q'])
        pi = np.array(case['root_distn'])
        if 'raises' in case:
            with pytest.raises(getattr(S._util, case['raises'])):
                S._mjp_dense.get_likelihood(T, allowed, root, n, root_distn=pi, Q_default=Q)
            continue
        lk = S._mjp_dense.get_likelihood(T, allowed, root, n, root_distn=pi, Q_default=Q)
        assert_allclose(lk, case['likelihood'], rtol=1e-10)
        dwell, rootp, trans = S._mjp_dense.get_expected_history_statistics(
            T, allowed, root, n, root_distn=pi, Q_default=Q)
        assert isinstance(dwell, dict) and isinstance(trans, nx.DiGraph)
        assert_allclose([dwell.get(s, 0.0) for s in range(n)], case['dwell'], rtol=1e-9, atol=1e-12)
        assert_allclose(rootp, case['root_post'], rtol=1e-10, atol=1e-14)
        tr = np.zeros((n, n))
        for a, b, d in trans.edges(data=True):
            tr[a, b] = d['weight']
        assert_allclose(tr, np.array(case['trans']), rtol=1e-9, atol=1e-12)
        T_aug = S._mjp_dense.get_expm_augmented_tree(T, root, Q_default=Q)
        for a, b, P in case['P']:
            assert_allclose(T_aug[a][b]['P'], np.array(P), rtol=0, atol=1e-13)
        pmap = S._mcy_dense.get_node_to_pmap(T_aug, root, n, node_to_allowed_states=allowed)
        for v, p in case['pmap'].items():
            assert_allclose(pmap[int(v)], p, rtol=1e-10, atol=1e-300)
        distn = S._mc0_dense.get_node_to_distn(T_aug, root, pmap, n, root_distn=pi)
        for v, d in case['node_distn'].items():
            assert_allclose(distn[int(v)], d, rtol=1e-10, atol=1e-14)
        TJ = S._mc0_dense.get_joint_endpoint_distn(T_aug, root, pmap, distn, n)
        for a, b, J in case['joint']:
            assert_allclose(TJ[a][b]['J'], np.array(J), rtol=1e-10, atol=1e-14)
        pm2, dn2, ej2 = S._mcy_dense.kitchen_sink(T_aug, root, n, node_to_allowed_states=allowed,
                                                  root_distn=pi)
        for a, b, J in case['joint']:
            assert_allclose(ej2[(a, b)], np.array(J), rtol=1e-10, atol=1e-14)


def test_sparse_api_matches_reference_fixtures(S):
    g = load_golden('mjp_sparse.json')
    for case in g['cases']:
        T, root = case_tree(case)
        Q = nx.DiGraph()
        for a, b, w in case['Q']:
            Q.add_edge(a, b, weight=w)
        distn = dict((int(k), v) for k, v in case['root_distn'].items())
        allowed = case_allowed(case)
        if 'raises' in case:
            with pytest.raises(getattr(S._util, case['raises'])):
                S._mjp.get_likelihood(T, allowed, root, root_distn=distn, Q_default=Q)
            continue
        lk = S._mjp.get_likelihood(T, allowed, root, root_distn=distn, Q_default=Q)
        assert_allclose(lk, case['likelihood'], rtol=1e-10)
        dwell, rootp, trans = S._mjp.get_expected_history_statistics(
            T, allowed, root, root_distn=distn, Q_default=Q)
        for k, v in case['dwell'].items():
            assert_allclose(dwell[int(k)], v, rtol=1e-9, atol=1e-12)
        assert_equal(set(rootp), set(int(k) for k in case['root_post']))
        for k, v in case['root_post'].items():
            assert_allclose(rootp[int(k)], v, rtol=1e-10)
        assert_equal(set(trans.edges()), set((a, b) for a, b, w in case['trans']))
        for a, b, w in case['trans']:
            assert_allclose(trans[a][b]['weight'], w, rtol=1e-9, atol=1e-12)
        T_aug = S._mjp.get_expm_augmented_tree(T, root, Q_default=Q)
        for a, b, entries in case['P']:
            P = T_aug[a][b]['P']
            assert_equal(set(P.edges()), set((x, y) for x, y, w in entries))
            for x, y, w in entries:
                assert_allclose(P[x][y]['weight'], w, rtol=0, atol=1e-13)
        pset = S._mcy.get_node_to_pset(T_aug, root, node_to_allowed_states=allowed)
        nset = S._mcy.get_node_to_set(T_aug, root, node_to_allowed_states=allowed)
        for v, s in case['node_to_pset'].items():
            assert_equal(pset[int(v)], set(s))
        for v, s in case['node_to_set'].items():
            assert_equal(nset[int(v)], set(s))
        pmap = S._mcy.get_node_to_pmap(T_aug, root, node_to_allowed_states=allowed)
        for v, d in case['pmap'].items():
            assert_equal(set(pmap[int(v)]), set(int(k) for k in d))
            for k, x in d.items():
                assert_allclose(pmap[int(v)][int(k)], x, rtol=1e-10)
        nd = S._mc0.get_node_to_distn(T_aug, root, pmap, root_distn=distn)
        for v, d in case['node_distn'].items():
            for k, x in d.items():
                assert_allclose(nd[int(v)].get(int(k), 0.0), x, rtol=1e-10, atol=1e-14)
        emis = dict((int(v), dict((int(k), x) for k, x in d.items()))
                    for v, d in case['emissions'].items())
        zp = S._mcz.get_node_to_pmap(T_aug, root, node_to_state_to_likelihood=emis)
        for v, d in case['z_pmap'].items():
            for k, x in d.items():
                assert_allclose(zp[int(v)][int(k)], x, rtol=1e-10)


def test_get_total_rates(S):
    # raoteh/sampler/tests/test_mjp.py:30-50
    Q = nx.DiGraph()
    Q.add_weighted_edges_from([(0, 1, 1), (1, 0, 1), (1, 2, 1), (2, 1, 1)])
    assert_equal(S._mjp.get_total_rates(Q), {0: 1, 1: 2, 2: 1})
    Q_dense = S._density.rate_matrix_to_numpy_array(Q, nodelist=range(3))
    assert_allclose(S._mjp_dense.get_total_rates(Q_dense), [1, 2, 1])


def test_likelihoods_sum_to_one(S):
    # raoteh/sampler/tests/test_mjp.py:52-89
    rng = np.random.RandomState(0)
    T = nx.Graph()
    T.add_weighted_edges_from([(0, 1, 2.0), (0, 2, 3.0), (0, 3, 4.0)])
    nnodes, root, nstates = len(T), 0, 3
    distn = S._util.get_normalized_dict_distn(dict((i, rng.exponential()) for i in range(3)))
    Q = nx.DiGraph()
    for i in range(nstates):
        for j in range(nstates):
            if i != j:
                Q.add_edge(i, j, weight=rng.exponential())
    distn_dense = S._density.dict_to_numpy_array(distn, nodelist=range(nstates))
    Q_dense = S._density.rate_matrix_to_numpy_array(Q, nodelist=range(nstates))
    total = 0
    for assignment in product(range(nstates), repeat=nnodes):
        allowed = dict((n, {s}) for n, s in zip(range(nnodes), assignment))
        lk = S._mjp.get_likelihood(T, allowed, root, root_distn=distn, Q_default=Q)
        lk_dense = S._mjp_dense.get_likelihood(T, allowed, root, nstates,
                                               root_distn=distn_dense, Q_default=Q_dense)
        assert_allclose(lk, lk_dense)
        total += lk
    assert_allclose(total, 1)


def test_get_likelihood_root_invariance_and_marginalisation(S):
    # raoteh/sampler/tests/test_mjp.py:91-164
    T = nx.Graph()
    T.add_weighted_edges_from([(0, 1, 2.0), (0, 2, 3.0), (0, 3, 4.0), (1, 4, 5.0), (1, 5, 6.0)])
    nstates = 4
    distn = {0: 0.1, 1: 0.2, 2: 0.3, 3: 0.4}
    Q = nx.DiGraph()
    Q.add_weighted_edges_from([
        (0, 1, 1.0 * distn[1]), (1, 0, 1.0 * distn[0]), (1, 2, 2.0 * distn[2]),
        (2, 1, 2.0 * distn[1]), (2, 3, 1.0 * distn[3]), (3, 2, 1.0 * distn[2]),
        (3, 0, 2.0 * distn[0]), (0, 3, 2.0 * distn[3])])
    allowed = {0: set(range(4)), 1: set(range(4)), 2: {0}, 3: {1}, 4: {2}, 5: {3}}
    distn_dense = S._density.dict_to_numpy_array(distn, nodelist=range(nstates))
    Q_dense = S._density.rate_matrix_to_numpy_array(Q, nodelist=range(nstates))
    lks = []
    for root in range(6):
        lk = S._mjp.get_likelihood(T, allowed, root, root_distn=distn, Q_default=Q)
        lk_dense = S._mjp_dense.get_likelihood(T, allowed, root, nstates,
                                               root_distn=distn_dense, Q_default=Q_dense)
        assert_allclose(lk, lk_dense)
        lks.append(lk)
    assert_allclose(lks, lks[0])     # reversible Q with its stationary prior: root-invariant
    lk_m = 0
    for s0 in range(4):
        for s1 in range(4):
            nodemap = dict(allowed)
            nodemap[0] = {s0}
            nodemap[1] = {s1}
            lk_m += S._mjp.get_likelihood(T, nodemap, 3, root_distn=distn, Q_default=Q)
    assert_allclose(lks[0], lk_m)


def test_jukes_cantor_conditional_expectation(S):
    # raoteh/sampler/tests/test_mjp.py:166-240 (3 of the 16 end-state pairs, all roots)
    ce = S._conditional_expectation
    t = 0.5
    T = nx.Graph()
    for i, w in enumerate([0.1, 0.2, 0.3, 0.4]):
        T.add_edge(i, i + 1, weight=w * t)
    nstates = 4
    for a, b in ((0, 0), (0, 1), (2, 3)):
        allowed = {0: {a}, 1: set(range(4)), 2: set(range(4)), 3: set(range(4)), 4: {b}}
        Q = ce.get_jukes_cantor_rate_matrix(nstates)
        expected = [ce.get_jukes_cantor_interaction(a, b, i, i, t, nstates) /
                    ce.get_jukes_cantor_probability(a, b, t, nstates) for i in range(nstates)]
        Q_dense = S._density.rate_matrix_to_numpy_array(Q, nodelist=range(nstates))
        for root in T:
            dwell, init, trans = S._mjp.get_expected_history_statistics(T, allowed, root, Q_default=Q)
            assert_allclose([dwell[i] for i in range(nstates)], expected)
            dwell, init, trans = S._mjp_dense.get_expected_history_statistics(
                T, allowed, root, nstates, Q_default=Q_dense)
            assert_allclose([dwell[i] for i in range(nstates)], expected)


def test_code2x3_through_the_dense_api(S):
    """examples/code2x3/run.py main(): every likelihood and per-branch expectation."""
    g = load_golden('code2x3.json')
    for call in g['calls'][:14]:
        T, root = case_tree(call)
        allowed = case_allowed(call)
        n = call['nstates']
        Q, pi = np.array(call['Q']), np.array(call['root_distn'])
        if call['kind'] == 'likelihood':
            lk = S._mjp_dense.get_likelihood(T, allowed, root, n, root_distn=pi, Q_default=Q)
            assert_allclose(lk, call['out'], rtol=1e-10)
        else:
            E = None if call['E'] is None else np.array(call['E'])
            out = S._mjp_dense.get_expected_ntransitions(T, allowed, root, n, root_distn=pi,
                                                         Q_default=Q, E=E)
            for a, b, v in call['out']:
                assert_allclose(out[a, b], v, rtol=1e-9, atol=1e-12)


def test_gen_histories_invariants(S):
    # raoteh/sampler/tests/test_sampler.py:398-438
    T = nx.Graph()
    T.add_weighted_edges_from([(0, 12, 1.0), (0, 23, 2.0), (0, 33, 1.0), (23, 4, 0.5), (23, 5, 1.5)])
    Q = nx.DiGraph()
    Q.add_weighted_edges_from([(0, 1, 4), (0, 2, 2), (1, 0, 1), (1, 2, 2), (2, 1, 1), (2, 0, 2)])
    node_to_state = {12: 1, 33: 2, 4: 0, 5: 1}
    root_distn = {0: 0.25, 1: 0.5, 2: 0.25}
    total = T.size(weight='weight')
    n = 0
    for T_aug in S._sampler.gen_histories(T, Q, node_to_state, root=0, root_distn=root_distn,
                                          nhistories=12, seed=5):
        n += 1
        assert_allclose(T_aug.size(weight='weight'), total, rtol=1e-6)   # float32 event times
        for node, state in node_to_state.items():
            for nb in T_aug[node]:
                assert_equal(T_aug[node][nb]['state'], state)
        for node in T:
            assert len(set(T_aug[node][nb]['state'] for nb in T_aug[node])) == 1
        for node in set(T_aug) - set(T):
            assert node > max(T) and T_aug.degree(node) == 2
            s = [T_aug[node][nb]['state'] for nb in T_aug[node]]
            assert s[0] != s[1]
        dwell, root_state, trans = S._mjp.get_history_statistics(T_aug, root=0)
        assert_allclose(sum(dwell.values()), total, rtol=1e-6)
    assert n == 12


def test_gen_restricted_histories_input_validation(S):
    # raoteh/sampler/_sampler.py:329-343
    T = nx.Graph()
    T.add_edge(0, 1, weight=1.0)
    Q = nx.DiGraph()
    Q.add_weighted_edges_from([(0, 1, 1.0), (1, 0, 1.0)])
    ok = {0: {0}, 1: {1}}
    with pytest.raises(ValueError):
        next(S._sampler.gen_restricted_histories(T, Q, ok, 0, uniformization_factor=1))
    with pytest.raises(ValueError):
        next(S._sampler.gen_restricted_histories(T, nx.DiGraph(), ok, 0))
    Ql = Q.copy()
    Ql.add_edge(0, 0, weight=1.0)
    with pytest.raises(ValueError):
        next(S._sampler.gen_restricted_histories(T, Ql, ok, 0))
    with pytest.raises(ValueError):
        next(S._sampler.gen_restricted_histories(T, Q, {0: {0}, 7: {1}}, 0))
    with pytest.raises(ValueError):
        S._mjp_dense.get_likelihood(T, ok, 9, 2, Q_default=np.array([[-1., 1.], [1., -1.]]))


def test_tolerance_summary_matches_reference_fixture():
    """_tmjp_dense.get_tolerance_summary on primary trajectories sampled by
    the reference's own Rao-Teh sampler (tests/golden/tolerance_summary.json)."""
    from raoteh_b200.sampler import _tmjp_dense
    g = load_golden('tolerance_summary.json')
    Q_primary = np.array(g['Q_primary'])
    primary_to_part = dict((int(k), v) for k, v in g['primary_to_part'].items())
    for case in g['cases']:
        T_primary = nx.Graph()
        for a, b, w, s in case['edges']:
            T_primary.add_edge(a, b, weight=w, state=s)
        dd = None
        if case['disease'] is not None:
            dd = [dict((int(n), set(v)) for n, v in d.items()) for d in case['disease']]
        out = _tmjp_dense.get_tolerance_summary(primary_to_part, case['rate_on'], case['rate_off'],
                                                Q_primary, T_primary, case['root'], disease_data=dd)
        assert_allclose(out, case['out'], rtol=1e-9, atol=1e-12)
        contribs = _tmjp_dense.get_tolerance_ll_contribs(
            case['rate_on'], case['rate_off'], T_primary.size(weight='weight'), *out)
        assert all(np.isfinite(c) for c in contribs)


def test_reader_pipeline_matches_per_column_calls():
    """Text inputs (newick, codeml-style phylip, genetic code table) -> readers -> one batched
    evaluation == the per-column loop of examples/p53/p53.py:76-100 over the mirrored
    _mjp_dense.get_likelihood."""
    import io as _io
    from raoteh_b200 import io as rio, engine
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.sampler import _mjp_dense
    code_txt = ''.join('%d\t%s\t%s\n' % (i, r, c) for i, (r, c) in enumerate(
        # a connected set of codons (a 2 x 2 x 2 cube of single-nucleotide changes)
        [('ala', 'gct'), ('ala', 'gcc'), ('val', 'gtt'), ('val', 'gtc'), ('thr', 'act'),
         ('thr', 'acc'), ('ile', 'att'), ('ile', 'atc')])) + '8\tstop\ttaa\n'
    code = rio.read_genetic_code(_io.StringIO(code_txt))
    c2s, s2r, r2p, s2p = rio.codon_state_maps(code)
    Q, distn, _, _ = rio.mg94_from_genetic_code(0.25, 0.3, 0.25, 0.2, 2.5, 0.4, code,
                                                target_expected_rate=1.0)
    T, root, leaves = rio.read_newick(_io.StringIO('((Has:0.3,Ptr:0.2):0.1,(Mmu:0.25,Rno:0.4):0.15,Bta:0.5);'))
    rng = np.random.default_rng(3)
    codons = [c for s, r, c in code]
    names = [n for _, n in leaves]
    n_cols = 40
    seqs = {}
    for nme in names:
        col = [codons[k] for k in rng.integers(0, len(codons), n_cols)]
        if nme == 'Bta':
            col[5] = '---'
        seqs[nme] = col
    phy = '%d %d\n\n' % (len(names), 3 * n_cols) + ''.join('%s  %s\n\n' % (k, ''.join(v)) for k, v in seqs.items())
    rows = list(rio.read_phylip(_io.StringIO(phy), ntaxa=5, ncodons=n_cols))
    sched = TreeSchedule.from_nx(T, root)
    codes, leaf_nodes = rio.alignment_to_codes(rows, leaves, sched, c2s)
    mjp = engine.TreeMJP(sched, Q, root_distn=distn)
    obs = engine.Observations.from_leaf_codes(sched, codes, leaf_nodes)
    ll = mjp.log_likelihood(obs)['loglik'].cpu().numpy()
    name_to_leaf = dict((n, l) for l, n in leaves)
    for j in (0, 5, 17, 39):
        allowed = dict((v, set(range(8))) for v in T)
        for nme in names:
            cod = seqs[nme][j].upper()
            if cod in c2s:
                allowed[name_to_leaf[nme]] = {c2s[cod]}
        lk = _mjp_dense.get_likelihood(T, allowed, root, 8, root_distn=distn, Q_default=Q)
        assert_allclose(ll[j], np.log(lk), rtol=1e-10)


def test_linalg_special_cases_match_scipy():
    """raoteh/sampler/tests/test_expm.py:45-82 (the three restriction classes of the 3-state
    tolerance matrix: w > 0; w = 0, a != r; w = 0, a == r) for sparse_expm, plus the matching
    cases of simple_expm_frechet against scipy.linalg.expm_frechet, and the reachability pattern
    of sparse_expm_naive."""
    import scipy.linalg
    from raoteh_b200.sampler import _linalg
    from raoteh_b200.sampler._util import get_dense_rate_matrix
    rs = np.random.RandomState(1234)
    for t in np.logspace(-5, 5, 4, base=2):
        for kind in 'abc':
            a, w, r = rs.exponential(size=3)
            Q = nx.DiGraph()
            if kind == 'a':
                Q.add_weighted_edges_from([(0, 1, a), (1, 0, w), (1, 2, r)])
            elif kind == 'b':
                Q.add_weighted_edges_from([(0, 1, a), (1, 2, r)])
            else:
                Q.add_weighted_edges_from([(0, 1, a), (1, 2, a)])
            states, Qd = get_dense_rate_matrix(Q)
            want = scipy.linalg.expm(Qd * t)
            P = _linalg.sparse_expm(Q, t)
            got = np.zeros_like(want)
            for sa in states:
                for sb in states:
                    if P.has_edge(sa, sb):
                        got[sa, sb] = P[sa][sb]['weight']
            assert_allclose(got, want, rtol=1e-10, atol=1e-13)
            assert _linalg.expm_frechet_is_simple(Q)
            for (ai, bi, ci, di) in ((0, 0, 0, 0), (0, 1, 0, 1), (1, 1, 1, 1), (1, 0, 1, 0), (0, 1, 1, 1)):
                E = np.zeros((3, 3))
                E[ci, di] = 1.0
                K = scipy.linalg.expm_frechet(t * Qd, t * E, compute_expm=False)
                got = _linalg.simple_expm_frechet(Q, ai, bi, ci, di, t)
                assert_allclose(got, K[ai, bi], rtol=1e-9, atol=1e-13)
    # structural zeros of the generic path: 0 -> 1 -> 2, nothing leads back
    Q = nx.DiGraph()
    Q.add_weighted_edges_from([('x', 'y', 1.0), ('y', 'z', 2.0)])
    P = _linalg.sparse_expm(Q, 0.7)
    assert set(P.edges()) == {('x', 'x'), ('x', 'y'), ('x', 'z'), ('y', 'y'), ('y', 'z'), ('z', 'z')}
    assert_allclose(sum(P['x'][s]['weight'] for s in P['x']), 1.0, rtol=1e-12)


def _path_P(self_loops):
    P = nx.DiGraph()
    if self_loops:
        P.add_weighted_edges_from([(0, 0, 0.5), (1, 1, 0.5), (2, 2, 0.5), (3, 3, 0.5), (0, 1, 0.5),
                                   (1, 0, 0.25), (1, 2, 0.25), (2, 1, 0.25), (2, 3, 0.25), (3, 2, 0.5)])
    else:
        P.add_weighted_edges_from([(0, 1, 1.0), (1, 0, 0.5), (1, 2, 0.5), (2, 1, 0.5), (2, 3, 0.5),
                                   (3, 2, 1.0)])
    return P


def test_chain_state_samplers_deterministic_answers():
    """The known-answer tests of raoteh/sampler/tests/test_sample_mcx.py:21-149 (node states;
    short path, infeasible chain, separated regions) and :154-235 (edge states with event
    nodes): whenever only one assignment is feasible the sampler must return it, for every
    choice of root."""
    from raoteh_b200.sampler import _sample_mcx
    from raoteh_b200.sampler._util import StructuralZeroProb
    P = _path_P(False)
    T = nx.Graph()
    T.add_edges_from([(0, 1), (1, 2)])
    for root in T:
        with pytest.raises(StructuralZeroProb):
            _sample_mcx.resample_states(T, root, {0: 0, 2: 3}, P_default=P)
    uniform = {0: 0.25, 1: 0.25, 2: 0.25, 3: 0.25}
    for root in T:
        assert _sample_mcx.resample_states(T, root, {0: 0, 2: 2}, root_distn=uniform,
                                           P_default=P) == {0: 0, 1: 1, 2: 2}
        assert _sample_mcx.resample_states(T, root, {0: 3, 2: 1},
                                           root_distn={0: 0.1, 1: 0.2, 2: 0.3, 3: 0.4},
                                           P_default=P) == {0: 3, 1: 2, 2: 1}
        with pytest.raises(StructuralZeroProb):       # no transitions allowed at all
            _sample_mcx.resample_states(T, root, {0: 0, 2: 2}, root_distn=uniform,
                                        P_default=nx.DiGraph())
    T = nx.Graph()
    T.add_edges_from([(0, 10), (0, 20), (0, 30), (10, 11), (20, 21), (30, 31), (31, 32)])
    for root in T:
        got = _sample_mcx.resample_states(T, root, {0: 0, 11: 2, 21: 2, 32: 3}, root_distn=uniform,
                                          P_default=P)
        assert got == {0: 0, 10: 1, 11: 2, 20: 1, 21: 2, 30: 1, 31: 2, 32: 3}
    # edge states with event nodes (:154-235)
    P = _path_P(True)
    T = nx.Graph()
    T.add_weighted_edges_from([(0, 10, 1.0), (0, 20, 1.0), (0, 30, 1.0), (10, 11, 2.0), (11, 12, 2.0),
                               (12, 13, 2.0), (13, 14, 2.0), (14, 15, 2.0), (15, 16, 2.0),
                               (20, 21, 1.0), (21, 22, 1.0), (30, 31, 1.0), (31, 32, 1.0), (32, 33, 1.0)])
    node_to_state = {0: 0, 16: 0, 22: 2, 33: 3}
    event_nodes = {10, 20, 21, 30, 31, 32}
    for root in set(T) - event_nodes:
        T_aug = _sample_mcx.resample_edge_states(T, root, event_nodes, node_to_state=node_to_state,
                                                 root_distn=uniform, P_default=P)
        assert T.size() == T_aug.size()
        assert_allclose(T.size(weight='weight'), T_aug.size(weight='weight'))
        long_path = (10, 11, 12, 13, 14, 15, 16)
        for a, b in zip(long_path[:-1], long_path[1:]):
            assert T_aug[a][b]['state'] == 0
        assert [T_aug[a][b]['state'] for a, b in ((0, 20), (20, 21), (21, 22))] == [0, 1, 2]
        assert [T_aug[a][b]['state'] for a, b in ((0, 30), (30, 31), (31, 32), (32, 33))] == [0, 1, 2, 3]


def test_gen_histories_long_branch_birth_death():
    """raoteh/sampler/tests/test_sample_mjp.py:29-112 at the reference's own size: elapsed time
    2.0 with rates up to ~120 (about 490 candidate events on the single branch per sweep, above
    the 255 a branch holds on the device): the generator cuts the branch into pieces and joins
    them again.  Every history has an even number of excess events and keeps the tree length."""
    from raoteh_b200.sampler import _sampler
    from raoteh_b200.lowering import TreeSchedule, subdivide_long_branches
    T = nx.Graph()
    T.add_edge(0, 1, weight=2.0)
    Q = nx.DiGraph()
    for state in range(1, 51):
        if state - 1:
            Q.add_edge(state, state - 1, weight=(state - 1) * 1.5)
        if state < 50:
            Q.add_edge(state, state + 1, weight=state * 1.0)
    n = 0
    for traj in _sampler.gen_histories(T, Q, {0: 3, 1: 7}, root=0, root_distn=None, nhistories=25, seed=4):
        n += 1
        ntransitions = len(traj) - 2
        excess = ntransitions - (7 - 3)
        assert excess >= 0 and excess % 2 == 0
        assert_allclose(traj.size(weight='weight'), 2.0, rtol=1e-5)
        path = nx.shortest_path(traj, 0, 1)
        seq = [traj[a][b]['state'] for a, b in zip(path[:-1], path[1:])]
        assert seq[0] == 3 and seq[-1] == 7
        assert all(abs(x - y) == 1 for x, y in zip(seq[:-1], seq[1:]))
    assert n == 25
    # the subdivision itself
    s0 = TreeSchedule(np.array([-1, 0, 0, 2], dtype=np.int32), np.array([0.0, 1.0, 0.2, 2.5]))
    s1, pieces, image = subdivide_long_branches(s0, 1.0)
    assert s1.n == 4 + 0 + 0 + 2 and [len(pieces[c]) for c in (1, 2, 3)] == [1, 1, 3]
    assert_allclose(s1.length.sum(), s0.length.sum())
    assert (s1.parent[1:] < np.arange(1, s1.n)).all()
    assert_allclose(sum(s1.length[v] for v in pieces[3]), 2.5)


def test_graphed_evaluation_matches_eager_and_follows_the_rate_matrix():
    """TreeMJP.expected_history_statistics_graphed: the CUDA-graph replay of the evaluation gives
    the eager result, follows set_rate_matrix (the replay recomputes P from the new Q), and a second
    Observations object gets its own graph."""
    import torch
    from raoteh_b200 import engine, synth
    from raoteh_b200.lowering import TreeSchedule
    cfg = synth.config_c2(n_sites=5000, n_leaves=16)
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'])
    obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'])
    obs2 = engine.Observations.from_leaf_codes(sched, np.ascontiguousarray(cfg['codes'][:, :1234]), cfg['leaves'])
    for scale in (1.0, 0.7, 1.9, 1.0):
        Q = cfg['Q'] * scale
        mjp.set_rate_matrix(Q)
        ref = mjp.expected_history_statistics(obs)
        want = {k: ref[k].clone() for k in ('loglik', 'dwell', 'trans', 'root_post_sum')}
        for o, n in ((obs, 5000), (obs2, 1234)):
            mjp.set_rate_matrix(Q)
            got = mjp.expected_history_statistics_graphed(o)
            if n == 5000:
                for k in want:
                    np.testing.assert_allclose(got[k].cpu().numpy(), want[k].cpu().numpy(), rtol=1e-12, atol=1e-14)
                np.testing.assert_allclose(float(got['stats'][0]), float(want['loglik'].sum()), rtol=1e-12)
            else:
                np.testing.assert_allclose(got['loglik'].cpu().numpy(), want['loglik'].cpu().numpy()[:1234], rtol=1e-12)
    assert not getattr(mjp, '_graph_refused', False)
