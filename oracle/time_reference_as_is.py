"""
TEST / MEASUREMENT INFRASTRUCTURE -- not part of the product path.

Times the UNMODIFIED reference (/root/reference through oracle/ref_shim.py; `pyfelscore` is the
stand-in of oracle/pyfelscore_standin.py, not the Cython original) the way its applications
call it: ONE site per call in a Python loop (examples/p53/p53.py:88-100).

    python oracle/time_reference_as_is.py            # writes profiles/r2_reference_as_is.json

/root/reference exists only in the build container, so this cannot run inside bench.py on the
GPU box; bench.py quotes the committed record (`cpu_baseline.reference_as_is`).  One core: the
reference is single-threaded.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402


def nx_tree(parent, length):
    import networkx as nx
    T = nx.Graph()
    for b in range(1, len(parent)):
        T.add_edge(int(parent[b]), b, weight=float(length[b]))
    return T


def main():
    ref_shim.load_reference()
    import networkx as nx
    from raoteh.sampler import _mjp_dense, _sampler
    from raoteh_b200 import synth
    out = dict(where='build container, 1 core', python=sys.version.split()[0],
               note='unmodified reference, one site per call; pyfelscore = pure-Python stand-in')
    # ---- C2: likelihood + expectations per site
    cfg = synth.config_c2(n_sites=64)
    T = nx_tree(cfg['parent'], cfg['length'])
    E = len(cfg['parent']) - 1
    S = 4

    def allowed(site):
        d = {}
        for i, v in enumerate(cfg['leaves']):
            k = int(cfg['codes'][i, site])
            d[int(v)] = set(range(S)) if k == 255 else {k}
        for v in range(len(cfg['parent'])):
            d.setdefault(v, set(range(S)))
        return d
    n = 12
    t0 = time.perf_counter()
    for site in range(n):
        _mjp_dense.get_likelihood(T, allowed(site), 0, S, root_distn=cfg['pi'], Q_default=cfg['Q'])
    t_ll = (time.perf_counter() - t0) / n
    n = 4
    t0 = time.perf_counter()
    for site in range(n):
        _mjp_dense.get_expected_history_statistics(T, allowed(site), 0, S, root_distn=cfg['pi'],
                                                   Q_default=cfg['Q'])
    t_ex = (time.perf_counter() - t0) / n
    out['c2'] = dict(loglik_s_per_site=t_ll, loglik_messages_per_sec=E / t_ll,
                     step_s_per_site=t_ll + t_ex, step_messages_per_sec=E / (t_ll + t_ex),
                     functions='_mjp_dense.get_likelihood + _mjp_dense.get_expected_history_statistics',
                     sample='12 / 4 of the C2 sites')
    # ---- C4: Rao-Teh sweeps of one site
    cfg = synth.config_c4(n_sites=4)
    T = nx_tree(cfg['parent'], cfg['length'])
    Q = nx.DiGraph()
    for a in range(4):
        for b in range(4):
            if a != b:
                Q.add_edge(a, b, weight=float(cfg['Q'][a, b]))
    node_to_state = dict((int(v), int(cfg['codes'][i, 0])) for i, v in enumerate(cfg['leaves']))
    distn = dict((s, float(p)) for s, p in enumerate(cfg['pi']))
    n = 20
    gen = _sampler.gen_histories(T, Q, node_to_state, root=0, root_distn=distn, nhistories=n + 2)
    next(gen)
    next(gen)
    t0 = time.perf_counter()
    for _ in gen:
        pass
    t_sw = (time.perf_counter() - t0) / n
    out['c4'] = dict(sweeps_per_sec=1.0 / t_sw, functions='_sampler.gen_histories',
                     sample='%d sweeps of one C4 trajectory' % n)
    path = os.path.join(ROOT, 'profiles', 'r2_reference_as_is.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
