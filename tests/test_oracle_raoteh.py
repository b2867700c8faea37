"""The numpy restatement of the Rao-Teh sweep (oracle/np_oracle.py) against the
closed-form expectations, so the CPU baseline of the sweep is itself checked."""
import numpy as np

from oracle import np_oracle


def test_oracle_sweep_mean_dwell_matches_closed_form():
    rng = np.random.default_rng(2)
    parent = np.array([-1, 0, 0, 2, 2], dtype=np.int32)
    length = np.array([0.0, 0.4, 0.3, 0.5, 0.2])
    S = 3
    Q = rng.exponential(1.0, size=(S, S))
    np.fill_diagonal(Q, 0)
    Q -= np.diag(Q.sum(axis=1))
    pi = np.array([0.2, 0.3, 0.5])
    leaves = np.array([1, 3, 4])
    codes = np.array([[0], [2], [1]], dtype=np.uint8)
    P = np_oracle.expm_edges(Q, length)
    obs = np_oracle.Obs('codes', S, 1, leaf_nodes=leaves, codes=codes)
    want = np_oracle.expected_history_statistics(parent, length, Q, P, obs, pi)
    omega, B, rates = np_oracle.uniformized(Q, 2.0)
    allowed = np.ones((5, S))
    for i, v in enumerate(leaves):
        allowed[v] = 0
        allowed[v, codes[i, 0]] = 1
    traj = np_oracle.raoteh_init(parent, length, B, allowed, pi, rng)
    dwell = np.zeros(S)
    n = 3000
    for i in range(n + 100):
        traj = np_oracle.raoteh_sweep(parent, length, B, rates, allowed, pi, traj, rng)
        if i >= 100:
            d, r, tr = np_oracle.history_statistics(parent, length, traj, S)
            dwell += d
    np.testing.assert_allclose(dwell / n, want['dwell'], rtol=0.06)
    np.testing.assert_allclose(dwell.sum() / n, length.sum(), rtol=1e-9)
