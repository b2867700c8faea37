"""CPU restatement of the tolerance Gibbs sweep (oracle/np_tmjp.py gibbs_sweep) against the
closed form of its target law (pinned to the reference's sampler in test_oracle_golden.py)."""
import numpy as np

from oracle import np_tmjp


def test_cpu_gibbs_sweep_matches_closed_form():
    pre = np.array([[0, 1, 1, 0, 0, 0], [1, 0, 0, 1, 0, 0], [1, 0, 0, 1, 1, 0],
                    [0, 1, 1, 0, 0, 1], [0, 0, 1, 0, 0, 1], [0, 0, 0, 1, 1, 0]], dtype=float)
    Q = pre - np.diag(pre.sum(axis=1))
    Q /= -np.dot(np.ones(6) / 6, np.diag(Q))
    part = np.array([0, 0, 1, 1, 2, 2])
    pi = np.ones(6) / 6
    parent = np.array([-1, 0, 1, 2, 2, 1])
    length = np.array([0.0, 0.5, 0.7, 0.4, 0.9, 0.6])
    obs = {3: 4, 4: 5, 5: 1}
    disease = [{5: {1}, 3: {0}}, {4: {0}}, {5: {0, 1}}]
    want = np_tmjp.expected_sampler_statistics(parent, length, Q, part, 3, pi, 0.7, 1.3, obs, disease)
    want = np.concatenate([want['prim_dwell'], want['prim_trans'].ravel(), want['tol'].ravel()])
    groups, burn, n = 10, 20, 150
    means = np.zeros((groups, len(want)))
    for g in range(groups):
        rng = np.random.default_rng(100 + g)
        prim, tols = np_tmjp.gibbs_init(parent, length, Q, part, 3, pi, 0.7, 1.3, obs, disease, rng)
        acc = np.zeros(len(want))
        for i in range(burn + n):
            prim, tols = np_tmjp.gibbs_sweep(parent, length, Q, part, 3, pi, 0.7, 1.3, obs, disease,
                                             prim, tols, rng)
            if i >= burn:
                d, t, tl = np_tmjp.sampled_statistics(parent, length, prim, tols, 6, 3)
                acc += np.concatenate([d, t.ravel(), tl.ravel()])
        means[g] = acc / n
    m = means.mean(axis=0)
    se = means.std(axis=0, ddof=1) / np.sqrt(groups)
    for w, mm, s in zip(want, m, se):
        if w == 0:
            assert mm == 0
        else:
            s = max(s, np.sqrt(abs(w) / (groups * n)))
            assert abs(mm - w) < 5 * s, (w, mm, s)
