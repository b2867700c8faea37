"""
TEST INFRASTRUCTURE -- container only (needs /root/reference).

Generates the committed golden fixtures under tests/golden/ by running the
UNMODIFIED reference through oracle/ref_shim.py:

  code2x3.json      every _mjp_dense.get_likelihood / extras.get_expected_ntransitions
                    call made by examples/code2x3/run.py main() (pure primary,
                    switching, blinking models; data levels L0/L1/L2), inputs
                    and outputs recorded by wrapping the two functions.  The
                    printed values are the reference's published golden vectors
                    (examples/code2x3/full-description.tex:165-370).
  mjp_random.json   random small trees / rate matrices / allowed-state maps:
                    _mjp_dense.get_likelihood, get_expected_history_statistics,
                    _mcy_dense.get_node_to_pmap, _mc0_dense.get_node_to_distn,
                    get_joint_endpoint_distn outputs (seeded).
  jukes_cantor.json the closed-form check of raoteh/sampler/tests/test_mjp.py:166-240.

usage: python oracle/gen_golden.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(HERE, '..', 'tests', 'golden')


def tree_to_json(T, root):
    import networkx as nx
    edges = [(int(a), int(b), float(T[a][b]['weight'])) for a, b in nx.bfs_edges(T, root)]
    return dict(root=int(root), edges=edges, nodes=[int(v) for v in T])


def allowed_to_json(d):
    return dict((str(int(k)), sorted(int(s) for s in v)) for k, v in d.items())


def gen_code2x3():
    ref_shim.load_reference()
    ex = os.path.join(ref_shim.REFERENCE_ROOT, 'examples', 'code2x3')
    sys.path.insert(0, ex)
    import extras
    import run
    from raoteh.sampler import _mjp_dense
    calls = []
    orig_lik = _mjp_dense.get_likelihood
    orig_exp = extras.get_expected_ntransitions

    def rec_lik(T, node_to_allowed_states, root, nstates, root_distn=None, Q_default=None):
        out = orig_lik(T, node_to_allowed_states, root, nstates,
                       root_distn=root_distn, Q_default=Q_default)
        calls.append(dict(kind='likelihood', tree=tree_to_json(T, root), nstates=int(nstates),
                          allowed=allowed_to_json(node_to_allowed_states),
                          root_distn=np.asarray(root_distn).tolist(),
                          Q=np.asarray(Q_default).tolist(), out=float(out)))
        return out

    def rec_exp(T, node_to_allowed_states, root, nstates, root_distn=None, Q_default=None, E=None):
        out = orig_exp(T, node_to_allowed_states, root, nstates,
                       root_distn=root_distn, Q_default=Q_default, E=E)
        calls.append(dict(kind='edge_expectations', tree=tree_to_json(T, root),
                          nstates=int(nstates), allowed=allowed_to_json(node_to_allowed_states),
                          root_distn=np.asarray(root_distn).tolist(),
                          Q=np.asarray(Q_default).tolist(),
                          E=None if E is None else np.asarray(E).tolist(),
                          out=[[int(a), int(b), float(v)] for (a, b), v in out.items()]))
        return out

    _mjp_dense.get_likelihood = rec_lik
    run._mjp_dense.get_likelihood = rec_lik
    extras.get_expected_ntransitions = rec_exp
    run.extras.get_expected_ntransitions = rec_exp
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        run.main()
    _mjp_dense.get_likelihood = orig_lik
    extras.get_expected_ntransitions = orig_exp
    with open(os.path.join(OUT, 'code2x3.json'), 'w') as f:
        json.dump(dict(source='examples/code2x3/run.py main() via oracle/ref_shim.py',
                       printed=buf.getvalue(), calls=calls), f)
    print('code2x3: %d calls recorded' % len(calls))


def random_tree(rng, n):
    import networkx as nx
    T = nx.Graph()
    for b in range(1, n):
        a = int(rng.integers(0, b))
        T.add_edge(a, b, weight=float(rng.exponential(0.3) + 0.01))
    return T


def gen_mjp_random():
    ref_shim.load_reference()
    import networkx as nx
    mjpd = ref_shim.ref_module('_mjp_dense')
    mcyd = ref_shim.ref_module('_mcy_dense')
    mc0d = ref_shim.ref_module('_mc0_dense')
    util = ref_shim.ref_module('_util')
    rng = np.random.default_rng(1234)
    cases = []
    for S, n, sparse in [(2, 3, False), (3, 5, False), (4, 7, False), (4, 9, True),
                         (5, 8, False), (6, 10, True), (8, 12, False), (11, 9, False),
                         (16, 8, True), (20, 10, False),
                         (3, 5, 'tri'), (4, 6, 'tri'), (6, 7, 'tri'), (12, 6, 'tri')]:
        for rep in range(3 if sparse != 'tri' else 6):
            T = random_tree(rng, n)
            root = int(rng.integers(0, n))
            Q = rng.exponential(1.0, size=(S, S))
            if sparse == 'tri':   # reducible: structural zeros, some sites infeasible
                Q = np.triu(Q, 1)
            elif sparse:
                Q *= rng.random((S, S)) < 0.4
                for i in range(S):        # keep the chain irreducible-ish
                    Q[i, (i + 1) % S] += 0.5
            if sparse == 'tri':
                pass
            np.fill_diagonal(Q, 0)
            Q -= np.diag(Q.sum(axis=1))
            pi = rng.dirichlet(np.ones(S))
            allowed = {}
            deg = dict(T.degree())
            for v in T:
                r = rng.random()
                if deg[v] == 1 and r < 0.7:
                    allowed[v] = {int(rng.integers(0, S))}
                elif r < 0.3:
                    k = int(rng.integers(1, S + 1))
                    allowed[v] = set(int(x) for x in rng.choice(S, size=k, replace=False))
                else:
                    allowed[v] = set(range(S))
            case = dict(tree=tree_to_json(T, root), nstates=S, Q=Q.tolist(), root_distn=pi.tolist(),
                        allowed=allowed_to_json(allowed))
            try:
                lk = mjpd.get_likelihood(T, allowed, root, S, root_distn=pi, Q_default=Q)
                case['likelihood'] = float(lk)
                dwell, rootp, trans = mjpd.get_expected_history_statistics(
                    T, allowed, root, S, root_distn=pi, Q_default=Q)
                case['dwell'] = [float(dwell.get(s, 0.0)) for s in range(S)]
                case['root_post'] = np.asarray(rootp).tolist()
                tr = np.zeros((S, S))
                for a, b, d in trans.edges(data=True):
                    tr[a, b] = d['weight']
                case['trans'] = tr.tolist()
                T_aug = mjpd.get_expm_augmented_tree(T, root, Q_default=Q)
                pmap = mcyd.get_node_to_pmap(T_aug, root, S, node_to_allowed_states=allowed)
                case['pmap'] = dict((str(int(v)), np.asarray(p).tolist()) for v, p in pmap.items())
                distn = mc0d.get_node_to_distn(T_aug, root, pmap, S, root_distn=pi)
                case['node_distn'] = dict((str(int(v)), np.asarray(p).tolist()) for v, p in distn.items())
                TJ = mc0d.get_joint_endpoint_distn(T_aug, root, pmap, distn, S)
                case['joint'] = [[int(a), int(b), np.asarray(TJ[a][b]['J']).tolist()]
                                 for a, b in nx.bfs_edges(T, root)]
                case['P'] = [[int(a), int(b), np.asarray(T_aug[a][b]['P']).tolist()]
                             for a, b in nx.bfs_edges(T, root)]
            except util.ZeroProbError as e:
                case['raises'] = type(e).__name__
            cases.append(case)
    with open(os.path.join(OUT, 'mjp_random.json'), 'w') as f:
        json.dump(dict(source='oracle/gen_golden.py gen_mjp_random, seed 1234', cases=cases), f)
    print('mjp_random: %d cases (%d raise)' % (len(cases), sum('raises' in c for c in cases)))


def gen_jukes_cantor():
    ref_shim.load_reference()
    ce = ref_shim.ref_module('_conditional_expectation')
    t = 0.5
    n = 4
    rows = []
    for a in range(n):
        for b in range(n):
            exp = [ce.get_jukes_cantor_interaction(a, b, i, i, t, n) /
                   ce.get_jukes_cantor_probability(a, b, t, n) for i in range(n)]
            rows.append(dict(a=a, b=b, dwell=[float(x) for x in exp]))
    with open(os.path.join(OUT, 'jukes_cantor.json'), 'w') as f:
        json.dump(dict(source='raoteh/sampler/tests/test_mjp.py:166-240 closed forms '
                       '(_conditional_expectation.py:25-46)', t=t, nstates=n,
                       path_weights=[0.1 * t, 0.2 * t, 0.3 * t, 0.4 * t], rows=rows), f)
    print('jukes_cantor: %d rows' % len(rows))


def gen_sparse():
    """Sparse (networkx / dict) API: _mjp, _mcy, _mc0, _mcz on random cases with
    non-contiguous state labels."""
    ref_shim.load_reference()
    import networkx as nx
    mjp = ref_shim.ref_module('_mjp')
    mcy = ref_shim.ref_module('_mcy')
    mc0 = ref_shim.ref_module('_mc0')
    mcz = ref_shim.ref_module('_mcz')
    util = ref_shim.ref_module('_util')
    rng = np.random.default_rng(4321)
    cases = []
    for S, n, density in [(3, 4, 1.0), (4, 6, 0.6), (5, 7, 0.5), (4, 8, 1.0), (6, 6, 0.4)]:
        for rep in range(4):
            labels = [10 * (i + 1) + (i % 3) for i in range(S)]
            T = random_tree(rng, n)
            root = int(rng.integers(0, n))
            Q = nx.DiGraph()
            for i in range(S):
                for j in range(S):
                    if i != j and (rng.random() < density or j == (i + 1) % S):
                        Q.add_edge(labels[i], labels[j], weight=float(rng.exponential(1.0)))
            w = rng.dirichlet(np.ones(S))
            distn = dict((labels[i], float(w[i])) for i in range(S) if rep % 2 == 0 or i != 0)
            allowed = {}
            deg = dict(T.degree())
            for v in T:
                r = rng.random()
                if deg[v] == 1 and r < 0.7:
                    allowed[v] = {labels[int(rng.integers(0, S))]}
                elif r < 0.3:
                    k = int(rng.integers(1, S + 1))
                    allowed[v] = set(labels[int(x)] for x in rng.choice(S, size=k, replace=False))
                else:
                    allowed[v] = set(labels)
            case = dict(tree=tree_to_json(T, root), labels=labels,
                        Q=[[a, b, d['weight']] for a, b, d in Q.edges(data=True)],
                        root_distn=dict((str(k), v) for k, v in distn.items()),
                        allowed=allowed_to_json(allowed))
            try:
                case['likelihood'] = float(mjp.get_likelihood(T, allowed, root, root_distn=distn, Q_default=Q))
                dwell, rootp, trans = mjp.get_expected_history_statistics(
                    T, allowed, root, root_distn=distn, Q_default=Q)
                case['dwell'] = dict((str(k), float(v)) for k, v in dwell.items())
                case['root_post'] = dict((str(k), float(v)) for k, v in rootp.items())
                case['trans'] = [[a, b, float(d['weight'])] for a, b, d in trans.edges(data=True)]
                T_aug = mjp.get_expm_augmented_tree(T, root, Q_default=Q)
                case['P'] = [[int(a), int(b), [[x, y, float(d['weight'])] for x, y, d in
                                               T_aug[a][b]['P'].edges(data=True)]]
                             for a, b in nx.bfs_edges(T, root)]
                case['node_to_pset'] = dict((str(v), sorted(s)) for v, s in mcy.get_node_to_pset(
                    T_aug, root, node_to_allowed_states=allowed).items())
                case['node_to_set'] = dict((str(v), sorted(s)) for v, s in mcy.get_node_to_set(
                    T_aug, root, node_to_allowed_states=allowed).items())
                pmap = mcy.get_node_to_pmap(T_aug, root, node_to_allowed_states=allowed)
                case['pmap'] = dict((str(v), dict((str(k), float(x)) for k, x in d.items()))
                                    for v, d in pmap.items())
                nd = mc0.get_node_to_distn(T_aug, root, pmap, root_distn=distn)
                case['node_distn'] = dict((str(v), dict((str(k), float(x)) for k, x in d.items()))
                                          for v, d in nd.items())
                # z-type emissions on the same tree
                emis = dict((v, dict((s, float(rng.random() + 0.05)) for s in allowed[v])) for v in T)
                zp = mcz.get_node_to_pmap(T_aug, root, node_to_state_to_likelihood=emis)
                case['emissions'] = dict((str(v), dict((str(k), x) for k, x in d.items()))
                                         for v, d in emis.items())
                case['z_pmap'] = dict((str(v), dict((str(k), float(x)) for k, x in d.items()))
                                      for v, d in zp.items())
            except util.ZeroProbError as e:
                case['raises'] = type(e).__name__
            cases.append(case)
    with open(os.path.join(OUT, 'mjp_sparse.json'), 'w') as f:
        json.dump(dict(source='oracle/gen_golden.py gen_sparse, seed 4321', cases=cases), f)
    print('mjp_sparse: %d cases (%d raise)' % (len(cases), sum('raises' in c for c in cases)))


def gen_tolerance():
    """_tmjp_dense.get_tolerance_summary on primary trajectories sampled by the reference's
    own Rao-Teh sampler (toy model of _tmjp.get_example_tolerance_process_info)."""
    ref_shim.load_reference()
    import networkx as nx
    import random
    tmjpd = ref_shim.ref_module('_tmjp_dense')
    tmjp = ref_shim.ref_module('_tmjp')
    mjp = ref_shim.ref_module('_mjp')
    sampler = ref_shim.ref_module('_sampler')
    np.random.seed(7)
    random.seed(7)
    nprimary = 6
    pre = np.array([[0, 1, 1, 0, 0, 0], [1, 0, 0, 1, 0, 0], [1, 0, 0, 1, 1, 0],
                    [0, 1, 1, 0, 0, 1], [0, 0, 1, 0, 0, 1], [0, 0, 0, 1, 1, 0]], dtype=float)
    Q_primary = pre - np.diag(pre.sum(axis=1))
    Q_primary /= -np.dot(np.ones(nprimary) / nprimary, np.diag(Q_primary))
    Q_nx = nx.DiGraph()
    for i in range(nprimary):
        for j in range(nprimary):
            if i != j and Q_primary[i, j]:
                Q_nx.add_edge(i, j, weight=float(Q_primary[i, j]))
    primary_to_part = {0: 0, 1: 0, 2: 1, 3: 1, 4: 2, 5: 2}
    T = nx.Graph()
    for a, b, w in ((0, 1, 0.5), (1, 2, 0.7), (2, 3, 0.4), (2, 4, 0.9), (1, 5, 0.6)):
        T.add_edge(a, b, weight=w)
    root = 0
    node_to_state = {0: 0, 3: 4, 4: 5, 5: 1}
    distn = dict((i, 1.0 / nprimary) for i in range(nprimary))
    cases = []
    disease = [{0: {1}}, {0: {0}}, {0: {1}}]
    for k, T_primary in enumerate(sampler.gen_histories(T, Q_nx, node_to_state, root=root,
                                                        root_distn=distn, nhistories=8)):
        for rate_on, rate_off, dd in ((1.0, 1.0, None), (0.3, 2.0, disease)):
            out = tmjpd.get_tolerance_summary(primary_to_part, rate_on, rate_off, Q_primary,
                                              T_primary, root, disease_data=dd)
            extra = {}
            if dd is None:
                # the compound-process log-likelihood with the tolerance histories integrated out
                # (sparse module: the dense twin's body still indexes Q_primary like a graph) and
                # the trajectory log-likelihood under the primary process itself
                ctm_s = tmjp.CompoundToleranceModel(Q_nx, distn, primary_to_part, rate_on, rate_off)
                extra['tol_ll'] = float(tmjp.get_tolerance_process_log_likelihood(ctm_s, T_primary, root))
                extra['traj_ll'] = float(mjp.get_trajectory_log_likelihood(T_primary, root, distn, Q_nx))
            cases.append(dict(extra, 
                edges=[[int(a), int(b), float(T_primary[a][b]['weight']), int(T_primary[a][b]['state'])]
                       for a, b in nx.bfs_edges(T_primary, root)],
                root=root, rate_on=rate_on, rate_off=rate_off,
                disease=None if dd is None else [dict((str(n), sorted(s)) for n, s in d.items()) for d in dd],
                out=[float(x) for x in out]))
    with open(os.path.join(OUT, 'tolerance_summary.json'), 'w') as f:
        json.dump(dict(source='oracle/gen_golden.py gen_tolerance (reference _tmjp_dense.get_tolerance_summary)',
                       Q_primary=Q_primary.tolist(),
                       primary_to_part=dict((str(k), v) for k, v in primary_to_part.items()),
                       cases=cases), f)
    print('tolerance_summary: %d cases' % len(cases))


def _history_edges(T_aug, root):
    import networkx as nx
    return [(a, b, float(T_aug[a][b]['weight']), int(T_aug[a][b]['state']))
            for a, b in nx.bfs_edges(T_aug, root)]


def gen_tmjp_moments(nhistories=3000, burn=100):
    """Moments of the reference's own blocked Gibbs sampler
    (_sample_tmjp_dense.gen_histories_v1) on the toy tolerance model, with and without
    disease data: per-state primary dwell, primary transition counts, and per class
    (root on, dwell on, gains, losses).  Batch means give the Monte-Carlo standard errors."""
    ref_shim.load_reference()
    import networkx as nx
    import random
    tmjpd = ref_shim.ref_module('_tmjp_dense')
    stm = ref_shim.ref_module('_sample_tmjp_dense')
    nprimary = 6
    pre = np.array([[0, 1, 1, 0, 0, 0], [1, 0, 0, 1, 0, 0], [1, 0, 0, 1, 1, 0],
                    [0, 1, 1, 0, 0, 1], [0, 0, 1, 0, 0, 1], [0, 0, 0, 1, 1, 0]], dtype=float)
    Q_primary = pre - np.diag(pre.sum(axis=1))
    Q_primary /= -np.dot(np.ones(nprimary) / nprimary, np.diag(Q_primary))
    primary_distn = np.ones(nprimary) / nprimary
    primary_to_part = {0: 0, 1: 0, 2: 1, 3: 1, 4: 2, 5: 2}
    nparts = 3
    T = nx.Graph()
    tree_edges = ((0, 1, 0.5), (1, 2, 0.7), (2, 3, 0.4), (2, 4, 0.9), (1, 5, 0.6))
    for a, b, w in tree_edges:
        T.add_edge(a, b, weight=w)
    root = 0
    node_to_state = {3: 4, 4: 5, 5: 1}
    cases = []
    for name, rate_on, rate_off, dd in (
            ('plain', 0.7, 1.3, None),
            ('disease', 0.7, 1.3, [{5: {1}, 3: {0}}, {4: {0}}, {5: {0, 1}}])):
        np.random.seed(11)
        random.seed(11)
        ctm = tmjpd.CompoundToleranceModel(Q_primary, primary_distn, primary_to_part, rate_on, rate_off)
        rows = []
        for i, (T_prim, tol_trajs) in enumerate(stm.gen_histories_v1(
                ctm, T, root, node_to_state, disease_data=dd, nhistories=burn + nhistories)):
            if i < burn:
                continue
            dwell = np.zeros(nprimary)
            trans = np.zeros((nprimary, nprimary))
            for a, b in nx.bfs_edges(T_prim, root):
                dwell[T_prim[a][b]['state']] += T_prim[a][b]['weight']
            for v in T_prim:
                if v != root and T_prim.degree(v) == 2:
                    pred = [a for a, b in nx.bfs_edges(T_prim, root) if b == v][0]
                    succ = [b for a, b in nx.bfs_edges(T_prim, root) if a == v][0]
                    s0, s1 = T_prim[pred][v]['state'], T_prim[v][succ]['state']
                    if s0 != s1:
                        trans[s0, s1] += 1
            tol = np.zeros((nparts, 4))
            for c, tt in enumerate(tol_trajs):
                bfs = list(nx.bfs_edges(tt, root))
                first = [e for e in bfs if e[0] == root][0]
                tol[c, 0] = tt[first[0]][first[1]]['state']
                for a, b in bfs:
                    if tt[a][b]['state']:
                        tol[c, 1] += tt[a][b]['weight']
                pred_of = dict((b, a) for a, b in bfs)
                for a, b in bfs:
                    if a in pred_of:
                        s0, s1 = tt[pred_of[a]][a]['state'], tt[a][b]['state']
                        # count a change once per node: only along the first outgoing edge
                        if s0 != s1 and tt.degree(a) == 2:
                            tol[c, 2 if s1 else 3] += 1
            rows.append(np.concatenate([dwell, trans.ravel(), tol.ravel()]))
        rows = np.array(rows)
        nb = 30
        per = len(rows) // nb
        bm = rows[:nb * per].reshape(nb, per, -1).mean(axis=1)
        cases.append(dict(name=name, rate_on=rate_on, rate_off=rate_off,
                          disease=None if dd is None else [dict((str(n), sorted(s)) for n, s in d.items()) for d in dd],
                          nhistories=int(len(rows)), mean=rows.mean(axis=0).tolist(),
                          sem=(bm.std(axis=0, ddof=1) / np.sqrt(nb)).tolist()))
        print(name, 'done', len(rows))
    with open(os.path.join(OUT, 'tmjp_v1_moments.json'), 'w') as f:
        json.dump(dict(source='oracle/gen_golden.py gen_tmjp_moments (reference '
                              '_sample_tmjp_dense.gen_histories_v1, seed 11)',
                       layout='mean/sem = [dwell[6], trans[6*6], tol[3*4] = (root on, dwell on, gains, losses)]',
                       Q_primary=Q_primary.tolist(), primary_distn=primary_distn.tolist(),
                       primary_to_part=dict((str(k), v) for k, v in primary_to_part.items()),
                       tree_edges=[list(e) for e in tree_edges], root=root,
                       node_to_state=dict((str(k), v) for k, v in node_to_state.items()),
                       cases=cases), f)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == 'tolerance':
        gen_tolerance()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'tmjp':
        gen_tmjp_moments(int(sys.argv[2]) if len(sys.argv) > 2 else 3000)
        sys.exit(0)
    gen_tolerance()
    gen_sparse()
    gen_code2x3()
    gen_mjp_random()
    gen_jukes_cantor()
