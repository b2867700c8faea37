"""
TEST INFRASTRUCTURE -- not part of the product path.

Import the UNMODIFIED reference (/root/reference, Python-2 era, networkx 1.x,
numpy.testing.Tester, un-vendored pyfelscore) under Python 3.12 / numpy 2 /
networkx 3.  Used ONLY in the build container by oracle/gen_golden.py and by
CPU tests that validate the numpy restatement; /root/reference does not exist
on the GPU box, so nothing in `-m gpu` tests, smoke() or bench.py calls this.

Patches, applied before `import raoteh`:
  1. numpy.testing.{Tester, run_module_suite, decorators}  (raoteh/__init__.py:10-12)
  2. nx.to_numpy_matrix  (raoteh/sampler/_density.py:52)
  3. nx.all_pairs_shortest_path_length must return a dict
     (raoteh/sampler/_linalg.py:83-89; otherwise sparse_expm_naive silently
     returns an EMPTY matrix)
  4. DegreeView.items()  (raoteh/sampler/_graph_transform.py:171, _util.py:187)
  5. pyfelscore stand-in (oracle/pyfelscore_standin.py)
Nothing is written into /root/reference (sys.dont_write_bytecode).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get('RAOTEH_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'raoteh', 'sampler'))


_loaded = False


def load_reference():
    """Return the imported ``raoteh.sampler`` package of the reference."""
    global _loaded
    if not reference_available():
        raise RuntimeError('reference not present at %s' % REFERENCE_ROOT)
    if not _loaded:
        sys.dont_write_bytecode = True
        import numpy as np
        import numpy.testing as npt
        import networkx as nx

        # 1. numpy.testing stubs
        class _Tester(object):
            def __init__(self, *a, **k):
                pass

            def test(self, *a, **k):
                return None

            def bench(self, *a, **k):
                return None
        if not hasattr(npt, 'Tester'):
            npt.Tester = _Tester
        if not hasattr(npt, 'run_module_suite'):
            npt.run_module_suite = lambda *a, **k: None
        if not hasattr(npt, 'decorators'):
            dec = types.ModuleType('numpy.testing.decorators')
            dec.slow = lambda f: f
            dec.skipif = lambda *a, **k: (lambda f: f)
            dec.knownfailureif = lambda *a, **k: (lambda f: f)
            npt.decorators = dec
            sys.modules['numpy.testing.decorators'] = dec
        if not hasattr(npt, 'TestCase'):
            import unittest
            npt.TestCase = unittest.TestCase

        # 2. nx.to_numpy_matrix  (callers take `.A`)
        if not hasattr(nx, 'to_numpy_matrix'):
            class _Mat(np.ndarray):
                @property
                def A(self):
                    return np.asarray(self)

            def to_numpy_matrix(G, nodelist=None, **kwargs):
                return nx.to_numpy_array(G, nodelist=nodelist,
                                         **kwargs).view(_Mat)
            nx.to_numpy_matrix = to_numpy_matrix

        # 3. all_pairs_shortest_path_length -> dict
        _apspl = nx.all_pairs_shortest_path_length
        if not getattr(_apspl, '_rt_patched', False):
            def apspl(G, *a, **k):
                return dict(_apspl(G, *a, **k))
            apspl._rt_patched = True
            nx.all_pairs_shortest_path_length = apspl

        # 4. DegreeView.items
        from networkx.classes import reportviews as rv
        for name in ('DegreeView', 'DiDegreeView', 'InDegreeView',
                     'OutDegreeView', 'MultiDegreeView', 'DiMultiDegreeView'):
            cls = getattr(rv, name, None)
            if cls is not None and not hasattr(cls, 'items'):
                cls.items = lambda self: iter(dict(self).items())

        # 5. pyfelscore stand-in
        here = os.path.dirname(os.path.abspath(__file__))
        if here not in sys.path:
            sys.path.insert(0, here)
        import pyfelscore_standin
        sys.modules['pyfelscore'] = pyfelscore_standin

        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _loaded = True
    import raoteh.sampler  # noqa: F401
    return sys.modules['raoteh.sampler']


def ref_module(name: str):
    """e.g. ref_module('_mjp_dense') -> raoteh.sampler._mjp_dense"""
    load_reference()
    import importlib
    return importlib.import_module('raoteh.sampler.' + name)
