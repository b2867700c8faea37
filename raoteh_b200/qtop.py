"""Spectral P(t) scheme for time-reversible rate matrices.

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
