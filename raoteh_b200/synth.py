"""
Synthetic workloads C2-C4 of SURVEY.md section 8(d): random trees, HKY85 and
GY94/MG94 codon rate matrices, forward-simulated leaf states.  Host-side
numpy only; used by bench.py and the parity tests (same seeded inputs go to
the CUDA path and to the oracle).

Model recipes follow the reference's examples, not its code:
HKY as in raoteh/sampler/tests/test_mjp.py:91-121 (rates proportional to the
target frequency), MG94 as in examples/p53/create_mg94.py:81-128 with the
parameter values of examples/p53/p53.py:22-27, forward simulation with the
semantics of raoteh/sampler/_sampler.py:163-235 (get_forward_sample).
"""
from __future__ import annotations

import numpy as np

MISSING = 255


def random_binary_tree(n_leaves, mean_length, rng):
    """Random rooted binary tree: repeatedly join two random active lineages.

    Returns (parent, length, leaves) with nodes in DFS preorder (root 0,
    parent[i] < i, parent[0] = -1); length[i] is the branch above node i
    (length[0] = 0), drawn Exp(mean_length).
    """
    n = 2 * n_leaves - 1
    # build bottom-up with temporary ids, then relabel in preorder
    children = {}
    active = list(range(n_leaves))
    nxt = n_leaves
    while len(active) > 1:
        i, j = rng.choice(len(active), size=2, replace=False)
        a, b = active[i], active[j]
        children[nxt] = (a, b)
        active = [x for k, x in enumerate(active) if k not in (i, j)]
        active.append(nxt)
        nxt += 1
    root = active[0]
    parent = np.full(n, -1, dtype=np.int32)
    order = []
    label = {}
    stack = [(root, -1)]
    while stack:
        v, p = stack.pop()
        label[v] = len(order)
        order.append(v)
        parent[label[v]] = p
        for c in reversed(children.get(v, ())):
            stack.append((c, label[v]))
    length = rng.exponential(mean_length, size=n)
    length[0] = 0.0
    is_leaf = np.ones(n, dtype=bool)
    is_leaf[parent[1:]] = False
    leaves = np.nonzero(is_leaf)[0].astype(np.int32)
    return parent, length, leaves


def hky85(pi=(0.1, 0.2, 0.3, 0.4), kappa=2.0):
    """HKY85 rate matrix over (A, C, G, T)-like states 0..3, transitions
    0<->2 and 1<->3, scaled so that -sum_s pi_s Q_ss = 1."""
    pi = np.asarray(pi, dtype=float)
    Q = np.zeros((4, 4))
    for a in range(4):
        for b in range(4):
            if a != b:
                Q[a, b] = pi[b] * (kappa if (a + b) % 2 == 0 else 1.0)
    Q -= np.diag(Q.sum(axis=1))
    Q /= -(pi * np.diag(Q)).sum()
    return Q, pi


_NT = 'TCAG'
_AA = ('FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG')


def universal_code():
    """61 sense codons of the universal genetic code -> (codons, residues)."""
    codons, residues = [], []
    k = 0
    for a in _NT:
        for b in _NT:
            for c in _NT:
                aa = _AA[k]
                k += 1
                if aa != '*':
                    codons.append(a + b + c)
                    residues.append(aa)
    return codons, residues


def mg94(nt_freqs=None, kappa=3.17632, omega=0.21925, expected_rate=1.0):
    """61-state MG94-style codon model (examples/p53/create_mg94.py:81-128,
    parameters examples/p53/p53.py:22-27).  Returns (Q, pi, residues)."""
    if nt_freqs is None:
        nt_freqs = dict(A=0.25039, C=0.30126, G=0.25952, T=0.18883)
    codons, residues = universal_code()
    S = len(codons)
    transitions = ('AG', 'GA', 'CT', 'TC')
    Q = np.zeros((S, S))
    for i, ca in enumerate(codons):
        for j, cb in enumerate(codons):
            diff = [(x, y) for x, y in zip(ca, cb) if x != y]
            if len(diff) != 1:
                continue
            x, y = diff[0]
            rate = nt_freqs[y]
            if x + y in transitions:
                rate *= kappa
            if residues[i] != residues[j]:
                rate *= omega
            Q[i, j] = rate
    pi = np.array([np.prod([nt_freqs[c] for c in cod]) for cod in codons])
    pi /= pi.sum()
    Q -= np.diag(Q.sum(axis=1))
    Q *= expected_rate / -(pi * np.diag(Q)).sum()
    return Q, pi, residues


def simulate_leaf_codes(parent, length, leaves, Q, pi, n_sites, rng,
                        missing_frac=0.0):
    """Forward-simulate states from pi at the root down the tree and return
    uint8 codes [n_leaves, n_sites] (site-minor).  Uses P=expm(Qt) per edge,
    which has the same law as the jump-by-jump simulation of
    raoteh/sampler/_sampler.py:163-235 at the nodes."""
    import scipy.linalg
    n = len(parent)
    S = Q.shape[0]
    states = np.empty((n, n_sites), dtype=np.uint8)
    states[0] = rng.choice(S, size=n_sites, p=pi)
    for b in range(1, n):
        P = scipy.linalg.expm(Q * length[b])
        P = np.maximum(P, 0)
        cdf = np.cumsum(P, axis=1)
        cdf /= cdf[:, -1:]
        u = rng.random(n_sites)
        ps = states[parent[b]]
        # number of cdf entries strictly below u, per parent state (same result as comparing
        # against the gathered cdf rows, without the [n_sites, S] temporary)
        out = np.empty(n_sites, dtype=np.uint8)
        for a in range(S):
            idx = np.nonzero(ps == a)[0]
            if len(idx):
                out[idx] = np.searchsorted(cdf[a], u[idx], side='left')
        states[b] = out
    codes = states[leaves].copy()
    if missing_frac > 0:
        codes[rng.random(codes.shape) < missing_frac] = MISSING
    return np.ascontiguousarray(codes)


def config_c2(n_sites=1_000_000, seed=20260201, n_leaves=32):
    """C2: 4-state HKY, 32-leaf random binary tree, 1 % missing leaf cells."""
    rng = np.random.default_rng(seed)
    parent, length, leaves = random_binary_tree(n_leaves, 0.1, rng)
    Q, pi = hky85()
    codes = simulate_leaf_codes(parent, length, leaves, Q, pi, n_sites, rng, 0.01)
    return dict(name='C2', parent=parent, length=length, leaves=leaves, Q=Q,
                pi=pi, codes=codes, S=4)


def config_c3(n_sites=100_000, seed=20260202, n_leaves=128):
    """C3: 61-state codon model, 128-leaf tree."""
    rng = np.random.default_rng(seed)
    parent, length, leaves = random_binary_tree(n_leaves, 0.05, rng)
    Q, pi, _ = mg94()
    codes = simulate_leaf_codes(parent, length, leaves, Q, pi, n_sites, rng, 0.0)
    return dict(name='C3', parent=parent, length=length, leaves=leaves, Q=Q,
                pi=pi, codes=codes, S=61)


def config_c4(n_sites=10_000, seed=20260204, n_leaves=64):
    """C4: Rao-Teh Gibbs, 4-state HKY on a 64-leaf tree."""
    rng = np.random.default_rng(seed)
    parent, length, leaves = random_binary_tree(n_leaves, 0.1, rng)
    Q, pi = hky85()
    codes = simulate_leaf_codes(parent, length, leaves, Q, pi, n_sites, rng, 0.0)
    return dict(name='C4', parent=parent, length=length, leaves=leaves, Q=Q,
                pi=pi, codes=codes, S=4)


# topology of the 25-taxon mammalian p53 tree (examples/p53/p53S.const.tree), every branch 0.1
_P53_TOPOLOGY = ((((((0, 1), 2), (((3, 4), 5), 6)), (7, 8)),
                  ((((((9, 10), 11), (12, 13)), 14), (15, 16)), (17, 18))),
                 (19, ((20, 21), ((22, 23), 24))))


def tree_from_nested(nested, branch_length):
    """Nested tuples of leaf labels -> (parent, length, leaves, leaf_labels) in DFS preorder."""
    parent, leaves, labels = [], [], []

    def walk(x, p):
        i = len(parent)
        parent.append(p)
        if isinstance(x, tuple):
            for ch in x:
                walk(ch, i)
        else:
            leaves.append(i)
            labels.append(x)
    walk(nested, -1)
    n = len(parent)
    length = np.full(n, float(branch_length))
    length[0] = 0.0
    return (np.asarray(parent, dtype=np.int32), length, np.asarray(leaves, dtype=np.int32), labels)


def tolerance_proposal(Q, part, p_on):
    """Primary proposal rate matrix: between-class rates scaled by P(tolerance on)
    (raoteh/sampler/_tmjp_dense.py:1081-1131)."""
    Qp = Q - np.diag(np.diag(Q))
    cross = part[:, None] != part[None, :]
    Qp[cross] *= p_on
    return Qp - np.diag(Qp.sum(axis=1))


def config_c5(n_sites=1_000_000, seed=20260205):
    """C5: p53-style tolerance MJP.  Primary = the C3 codon model with omega = 1
    (examples/p53/blink.py:113-120), 20 amino-acid tolerance classes, rate_on = 0.21925,
    rate_off = 0.78075 (:122-127), 25-taxon tree with all lengths 0.1; sites simulated forward
    under the primary proposal model; disease data = tolerance states of the first leaf
    ('Has') fully observed, Bernoulli(0.2 off) except the leaf codon's own class (:244-269)."""
    rng = np.random.default_rng(seed)
    parent, length, leaves, _ = tree_from_nested(_P53_TOPOLOGY, 0.1)
    Q, pi, residues = mg94(omega=1.0)
    aas = sorted(set(residues))
    part = np.array([aas.index(r) for r in residues], dtype=np.int64)
    rate_on, rate_off = 0.21925, 0.78075
    Qp = tolerance_proposal(Q, part, rate_on / (rate_on + rate_off))
    codes = simulate_leaf_codes(parent, length, leaves, Qp, pi, n_sites, rng, 0.0)
    n_parts = len(aas)
    tol = np.where(rng.random((n_parts, n_sites)) < 0.2, 1, 2).astype(np.uint8)   # bit0 off / bit1 on
    tol[part[codes[0]], np.arange(n_sites)] = 2
    return dict(name='C5', parent=parent, length=length, leaves=leaves, Q=Q, Q_proposal=Qp, pi=pi,
                codes=codes, S=61, part=part, n_parts=n_parts, rate_on=rate_on, rate_off=rate_off,
                tol_obs=tol[None], tol_obs_nodes=[int(leaves[0])])
