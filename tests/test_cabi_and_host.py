"""CPU-only checks: the C-ABI library loads and exports every symbol the header
declares; host lowering (tree programs) is consistent."""
import ctypes
import os

import numpy as np
import pytest

from raoteh_b200 import _native, lowering, synth


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_native.LIB_PATH):
        from raoteh_b200 import _build
        _build.build()
    L = ctypes.CDLL(_native.LIB_PATH)
    declared = _native.declared_symbols()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(L, name), name
    for name in declared:
        assert name in _native.EXPORTS, 'binding missing for ' + name
    assert _native.lib().rt_version() >= 100


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from raoteh_b200 import engine
    parent, length, leaves = synth.random_binary_tree(4, 0.1, np.random.default_rng(0))
    with pytest.raises(_native.NativeError):
        engine.TreeMJP(lowering.TreeSchedule(parent, length), synth.hky85()[0])


def test_product_never_imports_the_oracle():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'raoteh_b200')
    offenders = []
    for dp, dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith('.py'):
                text = open(os.path.join(dp, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M):
                    offenders.append(f)
    assert offenders == []


@pytest.mark.parametrize('n_leaves', [2, 3, 8, 33, 128])
def test_up_program_is_a_valid_postorder_schedule(n_leaves):
    rng = np.random.default_rng(n_leaves)
    parent, length, leaves = synth.random_binary_tree(n_leaves, 0.1, rng)
    sched = lowering.TreeSchedule(parent, length)
    obs_slot = np.full(sched.n, -1, dtype=np.int32)
    obs_slot[leaves] = np.arange(len(leaves))
    ops, n_slots = sched.up_program(obs_slot)
    live = {}
    stored = set()
    msgs = set()
    last_store = None
    for code, node, a, b in ops:
        fresh = bool(code & lowering.OP_FLAG_FRESH)
        code &= 0xff
        if code == lowering.OP_MSG_SLOT:
            assert live.get(a) == node            # the slot still holds this child
            assert b == sched.store_index[node]
            if fresh:
                assert last_store == node
            msgs.add(node)
        elif code in (lowering.OP_MSG_OBS, lowering.OP_MSG_ONES):
            assert sched.is_leaf[node]
            msgs.add(node)
        elif code == lowering.OP_STORE:
            for c in sched.children[node]:
                assert c in msgs
                if not sched.is_leaf[c]:
                    slot = [k for k, v in live.items() if v == c]
                    for k in slot:
                        del live[k]
            assert a not in live and a < n_slots
            live[a] = node
            stored.add(node)
            last_store = node
        elif code == lowering.OP_ROOT:
            assert node == 0
    assert msgs == set(range(1, sched.n))
    assert stored == set(sched.internal) - {0}
    assert n_slots <= max(1, int(np.ceil(np.log2(n_leaves))) + 1)
    edges, level_ptr = sched.down_program(obs_slot)
    assert len(edges) == sched.n - 1 and level_ptr[-1] == sched.n - 1
    seen = {0}
    for row in edges:
        assert sched.parent[row[0]] in seen
        seen.add(int(row[0]))


def test_from_nx_matches_reference_preorder_convention():
    import networkx as nx
    T = nx.Graph()
    T.add_weighted_edges_from([(0, 1, 0.5), (1, 2, 0.5), (2, 3, 0.5), (2, 4, 0.5), (1, 5, 0.5)])
    sched = lowering.TreeSchedule.from_nx(T, 0)
    assert sched.nodes == list(nx.dfs_preorder_nodes(T, 0))
    assert all(sched.parent[i] < i for i in range(1, sched.n))
    with pytest.raises(ValueError):
        lowering.TreeSchedule.from_nx(T, 99)


def test_compound_tolerance_model_is_consistent():
    """_tmjp_dense.CompoundToleranceModel.init_compound (raoteh/sampler/_tmjp_dense.py:84-179):
    rows of Q_compound sum to zero, the compound distribution is stationary (the primary
    process is reversible), only compatible states carry mass."""
    import numpy as np
    from raoteh_b200.sampler import _tmjp_dense
    pre = np.array([[0, 1, 1, 0, 0, 0], [1, 0, 0, 1, 0, 0], [1, 0, 0, 1, 1, 0],
                    [0, 1, 1, 0, 0, 1], [0, 0, 1, 0, 0, 1], [0, 0, 0, 1, 1, 0]], dtype=float)
    Q = pre - np.diag(pre.sum(axis=1))
    part = {0: 0, 1: 0, 2: 1, 3: 1, 4: 2, 5: 2}
    ctm = _tmjp_dense.CompoundToleranceModel(Q, np.ones(6) / 6, part, 0.7, 1.3)
    assert (ctm.nprimary, ctm.nparts, ctm.ncompound) == (6, 3, 48)
    ctm.init_compound()
    np.testing.assert_allclose(ctm.Q_compound.sum(axis=1), 0, atol=1e-12)
    np.testing.assert_allclose(ctm.compound_distn.sum(), 1)
    np.testing.assert_allclose(ctm.compound_distn @ ctm.Q_compound, 0, atol=1e-12)
    for i, (p, t) in enumerate(zip(ctm.compound_to_primary, ctm.compound_to_tolerances)):
        assert (ctm.compound_distn[i] > 0) == (t[part[p]] == 1)
    Qp = _tmjp_dense.get_primary_proposal_rate_matrix(Q, part, ctm.tolerance_distn)
    np.testing.assert_allclose(Qp.sum(axis=1), 0, atol=1e-12)
    assert Qp[0, 1] == Q[0, 1] and np.isclose(Qp[0, 2], Q[0, 2] * 0.35)


import pytest  # noqa: E402


@pytest.mark.reference
def test_compound_tolerance_model_matches_reference():
    import numpy as np
    from oracle import ref_shim
    ref_shim.load_reference()
    ref = ref_shim.ref_module('_tmjp_dense')
    from raoteh_b200.sampler import _tmjp_dense
    rng = np.random.default_rng(2)
    Q = rng.exponential(1, (5, 5)) * (rng.random((5, 5)) < 0.6)
    np.fill_diagonal(Q, 0)
    Q -= np.diag(Q.sum(axis=1))
    part = {0: 0, 1: 1, 2: 1, 3: 2, 4: 0}
    distn = rng.dirichlet(np.ones(5))
    a = ref.CompoundToleranceModel(Q, distn, part, 0.4, 1.7)
    b = _tmjp_dense.CompoundToleranceModel(Q, distn, part, 0.4, 1.7)
    a.init_compound()
    b.init_compound()
    np.testing.assert_allclose(b.Q_compound, a.Q_compound, rtol=1e-14)
    np.testing.assert_allclose(b.compound_distn, a.compound_distn, rtol=1e-14)
    assert list(b.compound_to_primary) == list(a.compound_to_primary)
    assert [tuple(t) for t in b.compound_to_tolerances] == [tuple(t) for t in a.compound_to_tolerances]
    np.testing.assert_allclose(
        _tmjp_dense.get_primary_proposal_rate_matrix(Q, part, b.tolerance_distn),
        ref.get_primary_proposal_rate_matrix(Q, part, a.tolerance_distn), rtol=1e-14)


def test_forward_sample_invariants():
    """_sampler.get_forward_sample (raoteh/sampler/_sampler.py:163-235): tree length preserved,
    every jump follows an edge of Q, long-run dwell fractions approach the stationary distribution."""
    import networkx as nx
    import numpy as np
    from raoteh_b200.sampler import _sampler, _mjp
    np.random.seed(3)
    Q = nx.DiGraph()
    Q.add_weighted_edges_from([(0, 1, 1.0), (1, 0, 2.0), (1, 2, 1.0), (2, 1, 0.5)])
    T = nx.Graph()
    T.add_weighted_edges_from([(0, 1, 40.0), (1, 2, 30.0), (1, 3, 30.0)])
    dwell = {0: 0.0, 1: 0.0, 2: 0.0}
    for rep in range(20):
        H = _sampler.get_forward_sample(T, Q, 0, {0: 0.5, 1: 0.25, 2: 0.25})
        np.testing.assert_allclose(H.size(weight='weight'), 100.0)
        assert set(T) <= set(H)
        for a, b in nx.bfs_edges(H, 0):
            dwell[H[a][b]['state']] += H[a][b]['weight']
        for v in H:
            if v not in T:
                (x, y) = list(H[v])
                sa, sb = H[v][x]['state'], H[v][y]['state']
                assert sa != sb and (Q.has_edge(sa, sb) or Q.has_edge(sb, sa))
        d, root_state, trans = _mjp.get_history_statistics(H, root=0)
        assert abs(sum(d.values()) - 100.0) < 1e-9
    tot = sum(dwell.values())
    # stationary distribution of this chain: pi = (2, 1, 2) / 5
    np.testing.assert_allclose([dwell[s] / tot for s in (0, 1, 2)], [0.4, 0.2, 0.4], atol=0.05)
