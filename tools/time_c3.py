"""C3 timings: log-lik only, up pass with stored partials, expectations (up + down + contraction)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule
n_sites = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
cfg = synth.config_c3(n_sites=n_sites)
sched = TreeSchedule(cfg['parent'], cfg['length'])
mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'])
obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'])
ll = torch.empty(n_sites, dtype=torch.float64, device='cuda')
st = torch.empty(n_sites, dtype=torch.int8, device='cuda')
mjp.transition_matrices()
ev = lambda: torch.cuda.Event(enable_timing=True)


def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = ev(), ev()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


print('loglik  mean %.3f min %.3f ms' % t(lambda: mjp.log_likelihood(obs, out=(ll, st))), flush=True)
print('finite', bool(torch.isfinite(ll).all()), 'sum', float(ll.sum()), flush=True)
print('up+store mean %.3f min %.3f ms' % t(lambda: mjp.log_likelihood(obs, keep_partials=True, out=(ll, st))), flush=True)
print('expectations mean %.3f min %.3f ms' % t(lambda: mjp.expected_history_statistics(obs), 3), flush=True)
r = mjp.expected_history_statistics(obs)
print('dwell sum / (len*N)', float(r['dwell'].sum()) / (cfg['length'].sum() * n_sites), flush=True)
