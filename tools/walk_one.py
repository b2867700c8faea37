"""One C2-size evaluation through the two-kernel path (for ncu on down_walk_kernel)."""
import sys
import torch
sys.path.insert(0, '.')
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule
cfg = synth.config_c2(n_sites=1_000_000)
sched = TreeSchedule(cfg['parent'], cfg['length'])
obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'])
mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'])
for _ in range(3):
    r = mjp.posterior(obs, want_node_distn=False)
torch.cuda.synchronize()
print('ok', float(r['loglik'].sum()))
