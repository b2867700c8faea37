"""One C3-size expectations call (up pass with stored partials + DMMA down pass), for ncu."""
import sys
import torch
sys.path.insert(0, '.')
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule
cfg = synth.config_c3(n_sites=100_000)
sched = TreeSchedule(cfg['parent'], cfg['length'])
obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'])
mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'])
for _ in range(2):
    r = mjp.expected_history_statistics(obs)
torch.cuda.synchronize()
print('ok', float(r['loglik'].sum()), float(r['trans'].sum()))
