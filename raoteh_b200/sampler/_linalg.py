"""
Special-cased linear algebra of the reference (raoteh/sampler/_linalg.py): the sparse
matrix exponential with structural zeros and the Frechet-derivative entries of the 3-state
tolerance rate matrix.  The reference dispatches to closed forms in pyfelscore for the
restricted matrix [[-a, a, 0], [w, -w-r, r], [0, 0, 0]] (:31-37, :107-118); here every case
goes through the same two kernels, rt_expm_batched and rt_frechet_contract.
"""
from __future__ import division, print_function, absolute_import

import ctypes  # noqa: F401

import networkx as nx
import numpy as np
import torch

from . import _sparse
from .. import _native

__all__ = []


def _expm(Q_dense, t):
    if not torch.cuda.is_available():
        raise _native.NativeError('raoteh_b200 needs a CUDA device; there is no CPU fallback')
    S = Q_dense.shape[0]
    dev = torch.device('cuda')
    Q = torch.from_numpy(np.ascontiguousarray(Q_dense[None])).to(dev)
    tt = torch.tensor([float(t)], dtype=torch.float64, device=dev)
    P = torch.empty((1, S, S), dtype=torch.float64, device=dev)
    rc = _native.lib().rt_expm_batched(Q.data_ptr(), None, tt.data_ptr(), 1, S, P.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
    _native.check(rc, 'rt_expm_batched')
    return P[0].cpu().numpy()


def _get_awr(Q):
    """raoteh/sampler/_linalg.py:14-29"""
    a = Q[0][1]['weight'] if Q.has_edge(0, 1) else 0
    w = Q[1][0]['weight'] if Q.has_edge(1, 0) else 0
    r = Q[1][2]['weight'] if Q.has_edge(1, 2) else 0
    return a, w, r


def _dense_with_diagonal(Q, states):
    Qd = _sparse.dense_matrix(Q, states)
    return Qd - np.diag(Qd.sum(axis=1))


def sparse_expm(Q, t):
    """raoteh/sampler/_linalg.py:31-40 -> nx.DiGraph of expm(tQ) with the reference's pattern."""
    edges = list(Q.edges())
    nonneg = all(Q[sa][sb]['weight'] >= 0 for sa, sb in edges)
    if nonneg and set(edges) <= {(0, 1), (1, 0), (1, 2)}:
        return sparse_expm_mmpp_block(Q, t)
    return sparse_expm_naive(Q, t)


def sparse_expm_mmpp_block(Q, t):
    """raoteh/sampler/_linalg.py:42-69: the 3-state tolerance matrix; entries that are
    structurally zero for the given (a, w, r) are left out."""
    a, w, r = _get_awr(Q)
    Qd = np.array([[-a, a, 0.0], [w, -w - r, r], [0.0, 0.0, 0.0]])
    P = _expm(Qd, t)
    P_nx = nx.DiGraph()
    P_nx.add_edge(0, 0, weight=P[0, 0])
    P_nx.add_edge(1, 1, weight=P[1, 1])
    P_nx.add_edge(2, 2, weight=1)
    if a:
        P_nx.add_edge(0, 1, weight=P[0, 1])
    if a and r:
        P_nx.add_edge(0, 2, weight=1 - P[0, 0] - P[0, 1])
    if w:
        P_nx.add_edge(1, 0, weight=P[1, 0])
    if r:
        P_nx.add_edge(1, 2, weight=1 - P[1, 0] - P[1, 1])
    return P_nx


def sparse_expm_naive(Q, t):
    """raoteh/sampler/_linalg.py:72-90: dense expm, entries kept where the end state is
    reachable from the start state in the digraph of Q."""
    states = sorted(Q)
    P = _expm(_dense_with_diagonal(Q, states), t)
    return _sparse.sparse_matrix(P, states, pattern=_sparse.reachability(Q, states))


def expm_frechet_is_simple(Q):
    """raoteh/sampler/_linalg.py:92-104"""
    if len(Q) > 3:
        return False
    allowed = ((0, 1), (1, 0), (1, 2))
    if Q.size() > len(allowed):
        return False
    if not (set(Q.edges()) <= set(allowed)):
        return False
    a, w, r = _get_awr(Q)
    return not (a < 0 or r < 0 or w < 0)


def simple_expm_frechet(Q, ai, bi, ci, di, t):
    """raoteh/sampler/_linalg.py:107-118: expm_frechet(tQ, t E_{ci,di})[ai, bi] of the 3-state
    tolerance matrix (pyfelscore.get_mmpp_frechet_all_positive / _diagonalizable_w_zero /
    _defective_w_zero in the reference; one rt_frechet_contract call here)."""
    a, w, r = _get_awr(Q)
    Qd = np.array([[-a, a, 0.0], [w, -w - r, r], [0.0, 0.0, 0.0]])
    dev = torch.device('cuda')
    Qt = torch.from_numpy(Qd[None].copy()).to(dev)
    W = torch.zeros((1, 3, 3), dtype=torch.float64, device=dev)
    W[0, ai, bi] = 1.0
    tt = torch.tensor([float(t)], dtype=torch.float64, device=dev)
    M = torch.empty((1, 3, 3), dtype=torch.float64, device=dev)
    # M = L(t Q^T, t W): M[c, d] = sum_ab W[a, b] expm_frechet(tQ, t E_cd)[a, b]
    rc = _native.lib().rt_frechet_contract(Qt.data_ptr(), None, tt.data_ptr(), W.data_ptr(), 1, 3,
                                           M.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _native.check(rc, 'rt_frechet_contract')
    return float(M[0, ci, di])
