"""
Single-site evaluation helpers behind the reference-shaped functions of
raoteh_b200.sampler: they lower one (tree, matrices, observation map) call to a
batch of ONE site and run the same CUDA kernels as the batched engine.
"""
from __future__ import division, print_function, absolute_import

import networkx as nx
import numpy as np
import torch

from .. import engine
from ..lowering import TreeSchedule, check_square_dense
from ._util import NumericalZeroProb


def sched_and_P(T, root, nstates, P_default=None, P_key='P'):
    """Preorder schedule and dense [n,S,S] matrices indexed by the child node
    (raoteh/sampler/_density.py:143-180)."""
    sched = TreeSchedule.from_nx(T, root)
    P = np.zeros((sched.n, nstates, nstates), dtype=np.float64)
    for i in range(1, sched.n):
        a, b = sched.nodes[sched.parent[i]], sched.nodes[i]
        M = T[a][b].get(P_key, P_default)
        check_square_dense(M)
        if M.shape[0] != nstates:
            raise ValueError('transition matrix shape does not match nstates')
        P[i] = M
    return sched, P


def mask_from_allowed(sched, node_to_allowed_states, nstates):
    """dict node -> allowed state indices => python-int bitmasks per preorder node."""
    full = (1 << nstates) - 1
    mask = [full] * sched.n
    if node_to_allowed_states is not None:
        for v, allowed in node_to_allowed_states.items():
            i = sched.node_index.get(v)
            if i is None:
                continue
            m = 0
            for s in allowed:
                if 0 <= s < nstates:
                    m |= 1 << int(s)
            mask[i] = m
    return np.array(mask, dtype=np.uint64)


def bits(mask_row, nstates):
    return np.array([(int(mask_row) >> s) & 1 for s in range(nstates)], dtype=float)


class Evaluation(object):
    """Everything the reference-shaped wrappers need for ONE site."""

    def __init__(self, sched, P, root_distn, nstates):
        self.sched = sched
        self.S = nstates
        self.P = P
        self.root_distn = None if root_distn is None else np.asarray(root_distn, dtype=np.float64)
        self.mjp = engine.TreeMJP(sched, np.zeros((nstates, nstates)), root_distn=self.root_distn, P=P)

    def support(self, mask, passes=3):
        """Structural support (rt_support_sets); returns the pruned uint64 masks [n]."""
        dev = torch.from_numpy(mask.view(np.int64).reshape(-1, 1).copy()).to(self.mjp.device)
        self.mjp.support_sets(dev, passes=passes)
        return dev.cpu().numpy().view(np.uint64).reshape(-1)

    def upward_masks(self, mask, keep=True):
        """Pruning with hard masks at every node; returns (loglik, status, pmap[n,S])."""
        obs = engine.Observations.from_masks(self.sched, mask.reshape(-1, 1), device=self.mjp.device)
        r = self.mjp.log_likelihood(obs, keep_partials=keep, want_exponents=keep)
        self.obs = obs
        self.up = r
        pmap = None
        if keep:
            part = r['partials'][:, :, 0].cpu().numpy()
            expo = r['exponents'][:, 0].cpu().numpy()
            pmap = np.zeros((self.sched.n, self.S))
            for i in range(self.sched.n):
                k = self.sched.store_index[i]
                if k >= 0:
                    pmap[i] = np.ldexp(part[k], int(expo[k]))
                else:
                    pmap[i] = bits(mask[i], self.S)
        return float(r['loglik'][0]), int(r['status'][0]), pmap

    def upward_dense(self, lik, keep=True):
        """Pruning with emission likelihoods lik[n,S] at every node (obs type z)."""
        nodes = np.arange(self.sched.n)
        obs = engine.Observations.from_dense(self.sched, lik[:, :, None], nodes, device=self.mjp.device)
        r = self.mjp.log_likelihood(obs, keep_partials=keep, want_exponents=keep)
        self.obs = obs
        self.up = r
        pmap = None
        if keep:
            part = r['partials'][:, :, 0].cpu().numpy()
            expo = r['exponents'][:, 0].cpu().numpy()
            pmap = np.zeros((self.sched.n, self.S))
            for i in range(self.sched.n):
                k = self.sched.store_index[i]
                pmap[i] = np.ldexp(part[k], int(expo[k])) if k >= 0 else lik[i]
        return float(r['loglik'][0]), int(r['status'][0]), pmap

    def downward(self, want_joint=True):
        """Posterior marginals of every node D[n,S] and joints J[n,S,S] (by child)."""
        post = self.mjp.posterior(self.obs)
        J, D = self.mjp.joint_distn(self.obs, post)
        self.post = post
        return D[:, 0, :].cpu().numpy(), J[:, 0, :, :].cpu().numpy()

    def downward_given_pmap(self, pmap):
        """Down pass for a caller-supplied node_to_pmap (pmap[n,S])."""
        sched, S = self.sched, self.S
        root_w = pmap[0] if self.root_distn is None else pmap[0] * self.root_distn
        if not root_w.sum():
            raise NumericalZeroProb('the denominator is zero')
        leaves = sched.leaves
        dev = self.mjp.device
        if len(leaves):
            obs = engine.Observations.from_dense(sched, pmap[leaves][:, :, None], leaves, device=dev)
        else:
            obs = engine.Observations.from_dense(sched, np.ones((1, S, 1)), [], device=dev)
        partials = torch.from_numpy(np.ascontiguousarray(pmap[sched.internal][:, :, None])).to(dev)
        post = self.mjp.posterior_given_partials(obs, partials)
        J, D = self.mjp.joint_distn(obs, post)
        return D[:, 0, :].cpu().numpy(), J[:, 0, :, :].cpu().numpy()

    def expectations(self, Q_edges, lengths):
        """Per-edge contraction matrices M[n,S,S] for rate matrices Q_edges[n,S,S]
        (row 0 unused): M_b[c,d] = sum_ab (J_b/P_b)[a,b] * expm_frechet(tQ, tE_cd)[a,b]."""
        dev = self.mjp.device
        mjp = self.mjp
        W = self.post['W']
        Qd = torch.from_numpy(np.ascontiguousarray(Q_edges)).to(dev)
        qi = torch.arange(self.sched.n, dtype=torch.int32, device=dev)
        t = torch.from_numpy(np.ascontiguousarray(lengths, dtype=np.float64)).to(dev)
        M = torch.empty((self.sched.n, self.S, self.S), dtype=torch.float64, device=dev)
        from .. import _native
        rc = _native.lib().rt_frechet_contract(Qd.data_ptr(), qi.data_ptr(), t.data_ptr(),
                                               W.data_ptr(), self.sched.n, self.S, M.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream)
        _native.check(rc, 'rt_frechet_contract')
        M = M.cpu().numpy()
        M[0] = 0.0
        return M


def expm_edges(sched, T, nstates, Q_default, Q_key='Q'):
    """P[b] = expm(Q_b t_b) on the GPU (rt_expm_batched) with per-edge rate matrices
    honoured (T[a][b]['Q'], raoteh/sampler/_mjp_dense.py:355).  Returns (P[n,S,S], Q_edges[n,S,S])."""
    Qs = np.zeros((sched.n, nstates, nstates), dtype=np.float64)
    for i in range(1, sched.n):
        a, b = sched.nodes[sched.parent[i]], sched.nodes[i]
        Q = T[a][b].get(Q_key, Q_default)
        check_square_dense(Q)
        Qs[i] = Q
    mjp = engine.TreeMJP(sched, Qs, q_index=np.arange(sched.n, dtype=np.int32))
    P = mjp.transition_matrices().cpu().numpy()
    P[0] = 0.0
    return P, Qs
