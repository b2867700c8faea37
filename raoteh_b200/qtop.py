"""Alternative P(t) schemes: spectral form for time-reversible rate matrices,
Sylvester form for the block-structured switching model, small-step lower bounds.

Host-side mirror of the reference's `examples/p53/qtop.py` for the spectral form: the
decomposition (one symmetric eigenproblem per rate matrix, qtop.py:126-148) stays on the host like
in the reference; what the reference then does once per branch -- `getp_spectral_v2(D, A, lam, B, t)`
(qtop.py:76-85) -- is one batched device call for all branches (`rt_expm_spectral`).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg


def pseudo_reciprocal(v):
    """qtop.py:104-107: 1/v with 0 -> 0."""
    v = np.asarray(v, dtype=float)
    with np.errstate(divide='ignore'):
        r = np.reciprocal(v)
    return np.where(v == 0, v, r)


def decompose_spectral(S, D):
    """qtop.py:126-138: Q = dot(S, diag(D)), S symmetric, D >= 0 -> (D, U, lam) with
    (lam, U) = eigh(diag(sqrt D) S diag(sqrt D))."""
    S = np.asarray(S, dtype=float)
    D = np.asarray(D, dtype=float)
    d = np.sqrt(D)
    lam, U = scipy.linalg.eigh(d[:, None] * S * d[None, :])
    return D, U, lam


def decompose_spectral_v2(S, D):
    """qtop.py:140-148: (A, lam, B) with A = diag(D^-1/2) U, B = U^T diag(D^1/2), so that
    P(t) = A diag(exp(t lam)) B."""
    D, U, lam = decompose_spectral(S, D)
    d = np.sqrt(D)
    A = pseudo_reciprocal(d)[:, None] * U
    B = U.T * d[None, :]
    return A, lam, B


def reconstruct_spectral_v2(A, lam, B):
    """qtop.py:283-288."""
    return np.dot(A * np.asarray(lam)[None, :], B)


def symmetric_factor(Q, D, rtol=1e-9):
    """S with Q = dot(S, diag(D)) for a time-reversible Q with stationary weights D
    (qtop.py:395-402); raises ValueError if Q D^-1 is not symmetric (detailed balance)."""
    Q = np.asarray(Q, dtype=float)
    D = np.asarray(D, dtype=float)
    S = Q * pseudo_reciprocal(D)[None, :]
    on = D > 0
    Son = S[np.ix_(on, on)]
    scale = np.abs(Son).max() if Son.size else 0.0
    if Son.size and not np.allclose(Son, Son.T, rtol=rtol, atol=rtol * scale):
        raise ValueError('rate matrix is not time-reversible with respect to D')
    return S


def getp_spectral_v2(D, A, lam, B, t, device='cuda'):
    """qtop.py:76-85 for a whole vector of branch lengths `t`: P [len(t), S, S] (torch, on `device`)."""
    import torch
    from . import _native
    dev = torch.device(device)
    if dev.type != 'cuda' or not torch.cuda.is_available():
        raise _native.NativeError('getp_spectral_v2 needs a CUDA device (there is no CPU fallback)')
    S = len(lam)
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    put = lambda x, dt=np.float64: torch.from_numpy(np.ascontiguousarray(x, dtype=dt)).to(dev)
    Ad, ld, Bd, td = put(A), put(lam), put(B), put(t)
    off = put(np.asarray(D) == 0, np.uint8)
    P = torch.empty((len(t), S, S), dtype=torch.float64, device=dev)
    rc = _native.lib().rt_expm_spectral(Ad.data_ptr(), ld.data_ptr(), Bd.data_ptr(), td.data_ptr(),
                                        off.data_ptr(), len(t), S, P.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream)
    _native.check(rc, 'rt_expm_spectral')
    return P


# ---------------------------------------------------------------------------------------
# Sylvester form of the switching model (examples/p53/qtop.py:28-55, 150-266, 290-332):
#     Q = [[S0 D0 - diag(L), diag(L)], [0, S1 D1]]
# both diagonal blocks time-reversible.  expm(tQ) = [[R0, R0 X - X R1], [0, R1]] with
# R0 = expm(t (S0 D0 - L)), R1 = expm(t S1 D1) from their symmetric eigendecompositions and X the
# solution of the Sylvester equation (S0 D0 - L) X - X (S1 D1) = diag(L).
# The decompositions stay on the host like in the reference (eigh, Schur, trsyl); what the
# reference does once per branch -- getp_sylvester_v2 -- is batched over all branches on the device.
# ---------------------------------------------------------------------------------------
def build_block_2x2(A):
    """qtop.py:93-102"""
    (M11, M12), (M21, M22) = A
    n = M11.shape[0]
    M = np.empty((2 * n, 2 * n))
    M[:n, :n], M[:n, n:], M[n:, :n], M[n:, n:] = M11, M12, M21, M22
    return M


def _sym_eig(S, D, shift=None):
    d = np.sqrt(np.asarray(D, dtype=float))
    H = d[:, None] * np.asarray(S, dtype=float) * d[None, :]
    if shift is not None:
        H = H - np.diag(shift)
    lam, U = scipy.linalg.eigh(H)
    return lam, pseudo_reciprocal(d)[:, None] * U, U.T * d[None, :]


def decompose_sylvester_v2(S0, S1, D0, D1, L):
    """qtop.py:150-198 -> (A0, B0, A1, B1, L, lam0, lam1, XQ)."""
    S0, S1 = np.asarray(S0, dtype=float), np.asarray(S1, dtype=float)
    D0, D1, L = (np.asarray(x, dtype=float) for x in (D0, D1, L))
    lam0, A0, B0 = _sym_eig(S0, D0, shift=L)
    lam1, A1, B1 = _sym_eig(S1, D1)
    # solve_sylvester(A, B, Q): A X + X B = Q
    XQ = scipy.linalg.solve_sylvester(S0 * D0[None, :] - np.diag(L), -(S1 * D1[None, :]), np.diag(L))
    return A0, B0, A1, B1, L, lam0, lam1, XQ


def partial_syl_decomp_v3(S1, D1):
    """qtop.py:204-224: the part of the decomposition that depends on the default process only
    (reused while the reference process changes, e.g. inside an optimiser loop)."""
    S1, D1 = np.asarray(S1, dtype=float), np.asarray(D1, dtype=float)
    lam1, A1, B1 = _sym_eig(S1, D1)
    schur_S1, schur_V1 = scipy.linalg.schur(-(S1 * D1[None, :]).T, output='real')
    return A1, B1, lam1, schur_S1, schur_V1


def full_syl_decomp_v3(S0, D0, L, A1, B1, lam1, schur_S1, schur_V1):
    """qtop.py:227-265: second stage; same result as decompose_sylvester_v2."""
    S0, D0, L = (np.asarray(x, dtype=float) for x in (S0, D0, L))
    lam0, A0, B0 = _sym_eig(S0, D0, shift=L)
    schur_R0, schur_U0 = scipy.linalg.schur(S0 * D0[None, :] - np.diag(L), output='real')
    schur_F = np.dot(schur_U0.T * L[None, :], schur_V1)
    trsyl, = scipy.linalg.get_lapack_funcs(('trsyl',), (schur_R0, schur_S1, schur_F))
    Y, scale, info = trsyl(schur_R0, schur_S1, schur_F, tranb='C')
    if info < 0:
        raise Exception('lapack trsyl fail')
    XQ = np.dot(np.dot(schur_U0, scale * Y), schur_V1.T)
    return A0, B0, A1, B1, L, lam0, lam1, XQ


def reconstruct_sylvester_v2(A0, B0, A1, B1, L, lam0, lam1, XQ):
    """qtop.py:318-332 (host): with lam = eigenvalues -> Q, with exp(t lam) -> expm(tQ)."""
    R11 = reconstruct_spectral_v2(A0, lam0, B0)
    R22 = reconstruct_spectral_v2(A1, lam1, B1)
    return build_block_2x2([[R11, np.dot(R11, XQ) - np.dot(XQ, R22)],
                            [np.zeros_like(R11), R22]])


def getp_sylvester_v2(D0, A0, B0, A1, B1, L, lam0, lam1, XQ, t, device='cuda'):
    """qtop.py:44-55 for a whole vector of branch lengths `t`: P [len(t), 2n, 2n] (torch, on
    `device`).  The two diagonal blocks are rt_expm_spectral calls (states with D0 == 0 get a unit
    diagonal, as the reference forces), the off-diagonal block R0 X - X R1 two batched products."""
    import torch
    n = len(lam0)
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    R0 = getp_spectral_v2(D0, A0, lam0, B0, t, device=device)
    R1 = getp_spectral_v2(np.ones(n), A1, lam1, B1, t, device=device)
    X = torch.from_numpy(np.ascontiguousarray(XQ, dtype=np.float64)).to(R0.device)
    P = torch.zeros((len(t), 2 * n, 2 * n), dtype=torch.float64, device=R0.device)
    P[:, :n, :n] = R0
    P[:, n:, n:] = R1
    P[:, :n, n:] = torch.matmul(R0, X) - torch.matmul(X, R1)
    return P


# ---------------------------------------------------------------------------------------
# Small-step lower bounds (examples/p53/liwen.py:43-110; pyfelscore.get_lb_transition_matrix)
# ---------------------------------------------------------------------------------------
def getp_lb(Q, t, device='cuda'):
    """liwen.py:48-82 for a vector of time steps `t`: entry (a, b) = probability of NO change
    (a == b) or of exactly ONE change a -> b in the interval -- a lower bound of expm(tQ).
    P [len(t), S, S] (torch, on `device`), one rt_lb_transition call."""
    import torch
    from . import _native
    dev = torch.device(device)
    if dev.type != 'cuda' or not torch.cuda.is_available():
        raise _native.NativeError('getp_lb needs a CUDA device (there is no CPU fallback)')
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    S = Q.shape[0]
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    Qd = torch.from_numpy(Q).to(dev)
    td = torch.from_numpy(t).to(dev)
    P = torch.empty((len(t), S, S), dtype=torch.float64, device=dev)
    rc = _native.lib().rt_lb_transition(Qd.data_ptr(), td.data_ptr(), len(t), S, P.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream)
    _native.check(rc, 'rt_lb_transition')
    return P


def getp_bigt_lb(Q, dt, t, device='cuda'):
    """liwen.py:84-87: cut t into n = ceil(t / dt) steps and raise the small-step bound to the
    n-th power (repeated squaring on the device)."""
    import torch
    n = max(1, int(np.ceil(t / dt)))
    return torch.linalg.matrix_power(getp_lb(Q, t / n, device=device)[0], n)


def getp_approx(Q, t):
    """liwen.py:94-99: I + tQ."""
    Q = np.asarray(Q, dtype=float)
    return np.eye(*Q.shape) + Q * t


def getp_bigt_approx(Q, dt, t):
    """liwen.py:101-107: (I + Q t/n)^n."""
    n = max(1, int(np.ceil(t / dt)))
    return np.linalg.matrix_power(getp_approx(Q, t / n), n)
