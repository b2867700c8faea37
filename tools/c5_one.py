"""A few blocked Gibbs sweeps of the C5 tolerance model on 50 000 sites (for ncu on tmjp_kernel)."""
import sys
import torch
sys.path.insert(0, '.')
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule
from raoteh_b200.tmjp import ToleranceChains
n = 50_000
cfg = synth.config_c5(n_sites=n)
sched = TreeSchedule(cfg['parent'], cfg['length'])
obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'])
ch = ToleranceChains(sched, cfg['Q'], cfg['pi'], dict(enumerate(cfg['part'])), cfg['rate_on'],
                     cfg['rate_off'], obs, n_chains=1, tol_obs=cfg['tol_obs'],
                     tol_obs_nodes=cfg['tol_obs_nodes'], cap_p=96, cap_t=48, seed=20260205)
ch.initialize()
ch.sweep(3, stats=False)
ch.sweep(4)
torch.cuda.synchronize()
print('ok')
