"""
Drop-in mirror of the reference's `raoteh.sampler` private modules for the hot
path: same module names, function names, argument
meaning, return types and exceptions; every numerical step runs in the CUDA
library.  Like the reference (raoteh/sampler/__init__.py:7-9 exports nothing
useful), callers import the private modules directly:

    from raoteh_b200.sampler import _mjp_dense, _mjp, _mcy, _mcy_dense, _mcz, \
        _mc0, _mc0_dense, _sampler, _sample_mjp, _util, _density
"""
__all__ = []
