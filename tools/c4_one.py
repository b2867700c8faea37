"""A few C4 sweeps (for ncu): 64 chains x 1e4 sites."""
import sys
import torch
sys.path.insert(0, '.')
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule
from raoteh_b200.raoteh import RaoTehChains
cfg = synth.config_c4(n_sites=10_000)
sched = TreeSchedule(cfg['parent'], cfg['length'])
obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'])
ch = RaoTehChains(sched, cfg['Q'], obs, n_chains=48, root_distn=cfg['pi'], seed=1, cap=112)
ch.sweep(20, stats=False)
ch.sweep(5)
torch.cuda.synchronize()
print('ok', float(ch.dwell_sum.sum()))
