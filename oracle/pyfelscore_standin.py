"""
TEST INFRASTRUCTURE -- not part of the product path.

Stand-in for the un-vendored Cython dependency ``pyfelscore`` (named only in
the reference's README.md:8-9; no version pinned anywhere).  It exists so
that the UNMODIFIED reference under /root/reference can be imported in the
build container to generate golden vectors (oracle/gen_golden.py) and to
validate the numpy restatement (oracle/np_oracle.py).

Every function follows the reference's own pure-Python twin, cited per
function; the call signatures are those at the reference's call sites
(SURVEY.md section 8b).  All outputs are written in place.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg


def _children(idx, ptr, a):
    return idx[ptr[a]:ptr[a + 1]]


def _normalized(w):
    # raoteh/sampler/_util.py:151-166 (get_normalized_ndarray_distn)
    tot = w.sum()
    if not tot:
        from raoteh.sampler._util import NumericalZeroProb
        raise NumericalZeroProb('the denominator is zero')
    return w / tot


# spec: raoteh/sampler/_mcy.py:397-470 (unaccelerated_get_node_to_pset)
def mcy_esd_get_node_to_pset(tree_idx, tree_ptr, esd_P, state_mask):
    n = state_mask.shape[0]
    for a in range(n - 1, -1, -1):
        for b in _children(tree_idx, tree_ptr, a):
            reach = ((esd_P[b] > 0) & (state_mask[b][None, :] != 0)).any(axis=1)
            state_mask[a] &= reach.astype(state_mask.dtype)


# spec: raoteh/sampler/_mc0.py:89-138 (get_node_to_set_unaccelerated)
def esd_get_node_to_set(tree_idx, tree_ptr, esd_P, state_mask):
    n = state_mask.shape[0]
    for a in range(n):
        for b in _children(tree_idx, tree_ptr, a):
            reach = ((esd_P[b] > 0) & (state_mask[a][:, None] != 0)).any(axis=0)
            state_mask[b] &= reach.astype(state_mask.dtype)


def _csr_to_bool(trans_idx, trans_ptr, nstates):
    B = np.zeros((nstates, nstates), dtype=bool)
    for s in range(nstates):
        B[s, trans_idx[trans_ptr[s]:trans_ptr[s + 1]]] = True
    return B


# same two passes with one shared sparsity pattern (call sites _mcy.py:158,168)
def mcy_get_node_to_pset(tree_idx, tree_ptr, trans_idx, trans_ptr, state_mask):
    n, S = state_mask.shape
    B = _csr_to_bool(trans_idx, trans_ptr, S)
    for a in range(n - 1, -1, -1):
        for b in _children(tree_idx, tree_ptr, a):
            reach = (B & (state_mask[b][None, :] != 0)).any(axis=1)
            state_mask[a] &= reach.astype(state_mask.dtype)


def get_node_to_set(tree_idx, tree_ptr, trans_idx, trans_ptr, state_mask,
                    tmp_mask):
    n, S = state_mask.shape
    B = _csr_to_bool(trans_idx, trans_ptr, S)
    for a in range(n):
        for b in _children(tree_idx, tree_ptr, a):
            reach = (B & (state_mask[a][:, None] != 0)).any(axis=0)
            state_mask[b] &= reach.astype(state_mask.dtype)


# spec: raoteh/sampler/_mcy.py:611-682 (unaccelerated_get_node_to_pmap)
def mcy_esd_get_node_to_pmap(tree_idx, tree_ptr, esd_P, state_mask,
                             subtree_probability):
    n, S = state_mask.shape
    for a in range(n - 1, -1, -1):
        v = (state_mask[a] != 0).astype(float)
        for b in _children(tree_idx, tree_ptr, a):
            v = v * esd_P[b].dot(subtree_probability[b])
        subtree_probability[a] = v


# spec: raoteh/sampler/_mc0_dense.py:400-489 (get_node_to_distn)
def mc0_esd_get_node_to_distn(tree_idx, tree_ptr, esd_P, root_distn,
                              subtree_probability, node_to_distn):
    n, S = subtree_probability.shape
    w = root_distn * subtree_probability[0]
    node_to_distn[0] = _normalized(w)
    for a in range(n):
        for b in _children(tree_idx, tree_ptr, a):
            d = np.zeros(S, dtype=float)
            for sa in range(S):
                pa = node_to_distn[a, sa]
                if pa:
                    sw = esd_P[b, sa] * subtree_probability[b]
                    d += pa * _normalized(sw)
            node_to_distn[b] = d


# spec: raoteh/sampler/_mc0_dense.py:217-270 (get_joint_endpoint_distn)
def mc0_esd_get_joint_endpoint_distn(tree_idx, tree_ptr, esd_P,
                                     subtree_probability, node_to_distn,
                                     joint):
    n, S = subtree_probability.shape
    joint[...] = 0
    for a in range(n):
        for b in _children(tree_idx, tree_ptr, a):
            for sa in range(S):
                pa = node_to_distn[a, sa]
                if pa:
                    sw = esd_P[b, sa] * subtree_probability[b]
                    joint[b, sa] = pa * _normalized(sw)


# 3-state tolerance matrix, raoteh/sampler/_linalg.py:14-29, _tmjp.py:883-890
def _Q3(a, w, r):
    return np.array([[-a, a, 0.0], [w, -w - r, r], [0.0, 0.0, 0.0]])


def get_mmpp_block(a, w, r, t):
    return scipy.linalg.expm(t * _Q3(a, w, r))


def get_mmpp_block_zero_off_rate(a, r, t):
    return scipy.linalg.expm(t * _Q3(a, 0.0, r))


# equivalence stated by raoteh/sampler/_mjp.py:541-556
def _frechet(a, w, r, t, ai, bi, ci, di):
    C = np.zeros((3, 3))
    C[ci, di] = 1.0
    K = scipy.linalg.expm_frechet(t * _Q3(a, w, r), t * C, compute_expm=False)
    return float(K[ai, bi])


def get_mmpp_frechet_all_positive(a, w, r, t, ai, bi, ci, di):
    return _frechet(a, w, r, t, ai, bi, ci, di)


def get_mmpp_frechet_diagonalizable_w_zero(a, r, t, ai, bi, ci, di):
    return _frechet(a, 0.0, r, t, ai, bi, ci, di)


def get_mmpp_frechet_defective_w_zero(a, t, ai, bi, ci, di):
    return _frechet(a, 0.0, a, t, ai, bi, ci, di)


# asserted by raoteh/sampler/tests/test_expm.py:36-42
def get_tolerance_rate_matrix(t, Q, P):
    P[...] = scipy.linalg.expm(t * Q)


# spec: raoteh/sampler/_mjp_dense.py:496-533, _tmjp.py:588-607
def get_tolerance_expectations(t, Q, P, J, dwell, trans):
    def W(c, d):
        C = np.zeros((3, 3))
        C[c, d] = 1.0
        K = scipy.linalg.expm_frechet(t * Q, t * C, compute_expm=False)
        tot = 0.0
        for i in range(3):
            for j in range(3):
                if J[i, j]:
                    tot += J[i, j] * K[i, j] / P[i, j]
        return tot
    dwell[0] += W(0, 0)
    dwell[1] += W(1, 1)
    trans[0, 1] += Q[0, 1] * W(0, 1)
    trans[1, 0] += Q[1, 0] * W(1, 0)
    return Q[1, 2] * W(1, 1)


# spec: raoteh/sampler/_tmjp.py:815-902 (get_inhomogeneous_mjp)
def tmjp_get_inhomogeneous_mjp(tree_idx, tree_ptr, edge_to_primary_state,
                               primary_to_part, Q_primary, rate_on, rate_off,
                               tolerance_class, allowed, Q_tol):
    n = allowed.shape[0]
    nprimary = Q_primary.shape[0]
    for a in range(n):
        for b in _children(tree_idx, tree_ptr, a):
            s = edge_to_primary_state[b]
            same = primary_to_part[s] == tolerance_class
            off = 0.0 if same else rate_off
            absorb = 0.0
            for s2 in range(nprimary):
                if s2 != s and primary_to_part[s2] == tolerance_class:
                    absorb += Q_primary[s, s2]
            Q_tol[b] = _Q3(rate_on, off, absorb)
            if same:
                allowed[a, 0] = 0
                allowed[b, 0] = 0


def get_lb_transition_matrix(t, Q, P):  # examples/p53/liwen.py:45 only
    raise NotImplementedError('out of scope (SURVEY.md section 8b)')
