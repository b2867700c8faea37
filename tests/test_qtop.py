"""Spectral P(t) scheme: host mirror of examples/p53/qtop.py (CPU tests, the
reference's own round-trip / expm tests qtop.py:437-475, 562-610 ported) and the batched device
reconstruction rt_expm_spectral (GPU tests)."""
import numpy as np
import pytest
import scipy.linalg

from raoteh_b200 import qtop


def random_reversible_rate_matrix(n, rng, off_states=()):
    """qtop.py:340-372 (random symmetric rates, random stationary weights, rows sum to zero)."""
    S = np.square(rng.standard_normal((n, n)))
    S = S + S.T
    np.fill_diagonal(S, 0)
    D = np.square(rng.standard_normal(n)) + 0.05
    off = list(off_states)
    S[off, :] = 0
    S[:, off] = 0
    D[off] = 0
    D /= D.sum()
    pre_Q = S * D[None, :]
    Q = pre_Q - np.diag(pre_Q.sum(axis=1))
    return Q, Q * qtop.pseudo_reciprocal(D)[None, :], D


@pytest.mark.parametrize('n', [4, 7, 20])
def test_spectral_v2_round_trip_and_expm(n):
    rng = np.random.default_rng(1234 + n)
    Q, S, D = random_reversible_rate_matrix(n, rng)
    np.testing.assert_allclose(S, S.T, atol=1e-12)
    np.testing.assert_allclose(S * D[None, :], Q, atol=1e-14)
    A, lam, B = qtop.decompose_spectral_v2(S, D)
    np.testing.assert_allclose(qtop.reconstruct_spectral_v2(A, lam, B), Q, atol=1e-12)
    for t in (0.0, 0.23, 3.0):
        P = qtop.reconstruct_spectral_v2(A, np.exp(t * lam), B)
        np.testing.assert_allclose(P, scipy.linalg.expm(Q * t), atol=1e-13)
    np.testing.assert_allclose(qtop.symmetric_factor(Q, D), S, atol=1e-14)


def test_symmetric_factor_rejects_irreversible_matrices():
    rng = np.random.default_rng(5)
    Q = rng.exponential(1.0, size=(5, 5))
    np.fill_diagonal(Q, 0)
    Q -= np.diag(Q.sum(axis=1))
    with pytest.raises(ValueError):
        qtop.symmetric_factor(Q, np.full(5, 0.2))


def test_hky_and_mg94_are_reversible_with_respect_to_their_root_distribution():
    from raoteh_b200 import synth
    cfg = synth.config_c2(n_sites=8)
    qtop.symmetric_factor(cfg['Q'], cfg['pi'])
    Q, pi, _ = synth.mg94()
    A, lam, B = qtop.decompose_spectral_v2(qtop.symmetric_factor(Q, pi), pi)
    np.testing.assert_allclose(qtop.reconstruct_spectral_v2(A, np.exp(0.3 * lam), B),
                               scipy.linalg.expm(0.3 * Q), atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize('n,off', [(4, ()), (5, (0, 2)), (20, ()), (61, ())])
def test_rt_expm_spectral_matches_expm(n, off):
    """getp_spectral_v2 for a vector of branch lengths vs scipy.linalg.expm (and, for states with
    D == 0, the forced unit diagonal of qtop.py:83-84)."""
    rng = np.random.default_rng(99 + n)
    if n == 61:
        from raoteh_b200 import synth
        Q, D, _ = synth.mg94()
        S = qtop.symmetric_factor(Q, D)
    else:
        Q, S, D = random_reversible_rate_matrix(n, rng, off)
    A, lam, B = qtop.decompose_spectral_v2(S, D)
    t = np.concatenate([[0.0], rng.exponential(0.3, size=40), [5.0]])
    P = qtop.getp_spectral_v2(D, A, lam, B, t).cpu().numpy()
    for i, ti in enumerate(t):
        want = scipy.linalg.expm(Q * ti)
        want[np.asarray(D) == 0, np.asarray(D) == 0] = 1
        np.testing.assert_allclose(P[i], want, rtol=1e-10, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize('S', [4, 61])
def test_tree_likelihood_with_spectral_scheme_matches_pade(S):
    from raoteh_b200 import engine, synth
    from raoteh_b200.lowering import TreeSchedule
    cfg = synth.config_c2(n_sites=3000, n_leaves=16) if S == 4 else synth.config_c3(n_sites=700)
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'])
    obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'])
    ref = mjp.expected_history_statistics(obs)
    ref = {k: ref[k].clone() for k in ('loglik', 'dwell', 'trans', 'M_edges')}
    P_pade = mjp.transition_matrices().clone()
    mjp.use_spectral(cfg['pi'])
    P_spec = mjp.transition_matrices()
    np.testing.assert_allclose(P_spec[1:].cpu().numpy(), P_pade[1:].cpu().numpy(), rtol=1e-9, atol=1e-13)
    got = mjp.expected_history_statistics(obs)
    np.testing.assert_allclose(got['loglik'].cpu().numpy(), ref['loglik'].cpu().numpy(), rtol=1e-10)
    np.testing.assert_allclose(got['dwell'].cpu().numpy(), ref['dwell'].cpu().numpy(), rtol=1e-9)
    np.testing.assert_allclose(got['trans'].cpu().numpy(), ref['trans'].cpu().numpy(), rtol=1e-9, atol=1e-12)
    # the per-edge Frechet contraction itself: eigenbasis form against the Pade block exponential
    scale = float(ref['M_edges'].abs().max())
    np.testing.assert_allclose(got['M_edges'].cpu().numpy(), ref['M_edges'].cpu().numpy(), rtol=1e-8,
                               atol=1e-11 * scale)
    mjp.use_spectral(None)
    np.testing.assert_allclose(mjp.transition_matrices().cpu().numpy(), P_pade.cpu().numpy(), rtol=0, atol=0)


def random_for_sylvester(n, off_states, rng):
    """qtop.py:375-434: default process (S1, D1), reference process (S0, D0) with off states,
    switching rates L (zero at the off states)."""
    Q1, S1, D1 = random_reversible_rate_matrix(n, rng)
    Q0, S0, D0 = random_reversible_rate_matrix(n, rng, off_states)
    L = np.square(rng.standard_normal(n))
    L[list(off_states)] = 0
    return S0, S1, D0, D1, L


def _switching_Q(S0, S1, D0, D1, L):
    n = len(L)
    return qtop.build_block_2x2([[S0 * D0[None, :] - np.diag(L), np.diag(L)],
                                 [np.zeros((n, n)), S1 * D1[None, :]]])


def test_sylvester_round_trips():
    """qtop.py:489-560: one-stage and two-stage decompositions reconstruct the switching-model
    rate matrix, and with exp(t lam) its matrix exponential."""
    rng = np.random.default_rng(1234)
    S0, S1, D0, D1, L = random_for_sylvester(5, [0, 2], rng)
    Q = _switching_Q(S0, S1, D0, D1, L)
    dec = qtop.decompose_sylvester_v2(S0, S1, D0, D1, L)
    np.testing.assert_allclose(qtop.reconstruct_sylvester_v2(*dec), Q, atol=1e-12)
    part = qtop.partial_syl_decomp_v3(S1, D1)
    dec3 = qtop.full_syl_decomp_v3(S0, D0, L, *part)
    np.testing.assert_allclose(qtop.reconstruct_sylvester_v2(*dec3), Q, atol=1e-12)
    A0, B0, A1, B1, L_, lam0, lam1, XQ = dec
    for t in (0.1, 1.7):
        P = qtop.reconstruct_sylvester_v2(A0, B0, A1, B1, L_, np.exp(t * lam0), np.exp(t * lam1), XQ)
        off = np.nonzero(D0 == 0)[0]
        P[off, off] = 1          # the states the reference process never enters (qtop.py:52-54)
        np.testing.assert_allclose(P, scipy.linalg.expm(Q * t), atol=1e-12)


@pytest.mark.gpu
def test_sylvester_device_reconstruction_matches_expm():
    """getp_sylvester_v2 batched over branch lengths (two rt_expm_spectral calls + the off-diagonal
    block) against scipy.linalg.expm of the 2n x 2n switching-model rate matrix; the states with
    D0 == 0 get the unit diagonal the reference forces (qtop.py:52-54)."""
    rng = np.random.default_rng(77)
    for n, off in ((5, [0, 2]), (20, []), (61, [3])):
        S0, S1, D0, D1, L = random_for_sylvester(n, off, rng)
        Q = _switching_Q(S0, S1, D0, D1, L)
        A0, B0, A1, B1, L_, lam0, lam1, XQ = qtop.decompose_sylvester_v2(S0, S1, D0, D1, L)
        ts = np.array([0.0, 0.05, 0.4, 2.0])
        P = qtop.getp_sylvester_v2(D0, A0, B0, A1, B1, L_, lam0, lam1, XQ, ts).cpu().numpy()
        for k, t in enumerate(ts):
            want = scipy.linalg.expm(Q * t)
            np.testing.assert_allclose(P[k], want, atol=1e-11)
            np.testing.assert_allclose(P[k].sum(axis=1), 1.0, atol=1e-11)


@pytest.mark.gpu
def test_small_step_lower_bound_matches_oracle():
    """rt_lb_transition (liwen.py:48-82) against the numpy restatement, incl. the equal-exit-rate
    branch; the bound is below expm entry by entry and tight for small steps; getp_bigt_lb
    approaches expm as dt -> 0."""
    from oracle import np_oracle
    from raoteh_b200 import synth
    Q, pi = synth.hky85()
    Qe = np.array([[-1.0, 1.0, 0.0], [0.5, -1.0, 0.5], [0.0, 2.0, -2.0]])      # ra == rb for (0, 1)
    for M in (Q, Qe, synth.mg94()[0]):
        ts = np.array([1e-4, 0.01, 0.3])
        P = qtop.getp_lb(M, ts).cpu().numpy()
        for k, t in enumerate(ts):
            # the literal formula of the reference (exp(-ra t) - exp(-rb t)) / (rb - ra) cancels for
            # small steps (relative error ~ 1e-16 / ((rb - ra) t)); the kernel uses expm1
            np.testing.assert_allclose(P[k], np_oracle.getp_lb(M, t), rtol=3e-16 / (t * 1e-3) + 1e-12, atol=1e-300)
            ra = -np.diag(M)
            d = ra[None, :] - ra[:, None]                       # rb - ra
            with np.errstate(divide='ignore', invalid='ignore'):
                f = np.where(d == 0, t, -np.expm1(-d * t) / d)
            stable = M * np.exp(-ra * t)[:, None] * f
            np.fill_diagonal(stable, np.exp(-ra * t))
            np.testing.assert_allclose(P[k], stable, rtol=1e-13, atol=1e-300)
            assert (P[k] <= scipy.linalg.expm(M * t) + 1e-15).all()
        np.testing.assert_allclose(P[0], scipy.linalg.expm(M * ts[0]), atol=1e-6)
    big = qtop.getp_bigt_lb(Q, 1e-3, 0.5).cpu().numpy()
    np.testing.assert_allclose(big, scipy.linalg.expm(Q * 0.5), atol=2e-3)
    assert (big <= scipy.linalg.expm(Q * 0.5) + 1e-12).all()
    np.testing.assert_allclose(qtop.getp_bigt_approx(Q, 1e-4, 0.5), scipy.linalg.expm(Q * 0.5), atol=1e-3)
