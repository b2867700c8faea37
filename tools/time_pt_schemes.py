"""P(t) for all branches after a rate-matrix update: Pade (rt_expm_batched) vs the spectral scheme
(host eigh + rt_expm_spectral), at C2 (4 states, 62 branches) and C3 (61 states, 254 branches)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule
dev = torch.device('cuda:0')
out = {}
for name, cfg in (('c2', synth.config_c2(n_sites=256)), ('c3', synth.config_c3(n_sites=256))):
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'], device=dev)
    res = {}
    for scheme in ('pade', 'spectral'):
        mjp.use_spectral(cfg['pi'] if scheme == 'spectral' else None)
        for _ in range(3):
            mjp.set_rate_matrix(cfg['Q']); mjp.transition_matrices()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 20
        for _ in range(n):
            mjp.set_rate_matrix(cfg['Q']); mjp.transition_matrices()
        torch.cuda.synchronize()
        res[scheme + '_ms_per_update'] = (time.perf_counter() - t0) / n * 1e3
    out[name] = res
print(json.dumps(out))
