"""
Exception classes and small helpers of the reference's `raoteh.sampler._util`
(raoteh/sampler/_util.py), part of the API contract.
"""
from __future__ import division, print_function, absolute_import

import numpy as np

__all__ = []


# raoteh/sampler/_util.py:14-21
class ZeroProbError(Exception):
    pass


class StructuralZeroProb(ZeroProbError):
    pass


class NumericalZeroProb(ZeroProbError):
    pass


def get_first_element(elements):
    for x in elements:
        return x


def get_dense_rate_matrix(Q_sparse):
    """raoteh/sampler/_util.py:27-53: (sorted states, dense Q with diagonal)."""
    from ..lowering import dense_rate_matrix
    return dense_rate_matrix(Q_sparse)


def get_unnormalized_dict_distn(d, prior=None):
    # raoteh/sampler/_util.py:88-101
    if d is None:
        raise ValueError('d is None')
    if not d:
        raise StructuralZeroProb('the main dict of weights is empty')
    if prior is None:
        return d
    if not prior:
        raise StructuralZeroProb('empty prior')
    states = set(d) & set(prior)
    if not states:
        raise StructuralZeroProb('empty intersection of main and prior')
    return dict((k, d[k] * prior[k]) for k in states)


def get_normalized_dict_distn(d, prior=None):
    # raoteh/sampler/_util.py:104-109
    dpost = get_unnormalized_dict_distn(d, prior)
    total_weight = sum(dpost.values())
    if not total_weight:
        raise NumericalZeroProb('the denominator is zero')
    return dict((k, v / total_weight) for k, v in dpost.items())


def get_unnormalized_ndarray_distn(d, prior=None, atol=1e-6):
    # raoteh/sampler/_util.py:112-148
    d_min = d.min()
    if d_min < -atol:
        raise ValueError('expected non-negative entries but found ' + str(d_min))
    if prior is None:
        return d
    prior_min = prior.min()
    if prior_min < -atol:
        raise ValueError('expected non-negative prior entries but found ' + str(prior_min))
    return d * prior


def get_normalized_ndarray_distn(d, prior=None):
    # raoteh/sampler/_util.py:151-166
    dpost = get_unnormalized_ndarray_distn(d, prior)
    total_weight = dpost.sum()
    if not total_weight:
        raise NumericalZeroProb('the denominator is zero')
    return dpost / total_weight


def get_arbitrary_tip(T, degrees=None):
    # raoteh/sampler/_util.py:169-189
    if degrees is None:
        degrees = dict(T.degree())
    return get_first_element(n for n, d in dict(degrees).items() if d == 1)


def _check_root(T, root):
    if root not in T:
        raise Exception('internal error: the root is not in the tree')
