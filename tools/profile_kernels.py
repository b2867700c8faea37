"""Runs each hot kernel once (for `ncu --set full`): K2 (codes / dense / store),
K4 down pass at C2 size; K3 (DMMA pruning) and K5 (DMMA down pass) at C3 size."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raoteh_b200 import engine, synth  # noqa: E402
from raoteh_b200.lowering import TreeSchedule  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
dev = torch.device('cuda:0')
if which in ('all', 'c2'):
    cfg = synth.config_c2(n_sites=1_000_000)
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'], device=dev)
    obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'], device=dev)
    mjp.log_likelihood(obs)
    S, N = 4, obs.n_sites
    codes = obs.data.long()
    lik = torch.zeros((len(sched.leaves), S, N), dtype=torch.float64, device=dev)
    lik.scatter_(1, codes.clamp(max=S - 1).unsqueeze(1), 1.0)
    lik[(codes == 255).unsqueeze(1).expand(-1, S, -1)] = 1.0
    dobs = engine.Observations(engine.OBS_DENSE, lik, obs.obs_slot, N)
    mjp.log_likelihood(dobs)
    r = mjp.expected_history_statistics(obs)
    print('c2 ok', float(r['loglik'].sum()))
if which in ('all', 'c3'):
    cfg = synth.config_c3(n_sites=100_000)
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'], device=dev)
    obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'], device=dev)
    r = mjp.log_likelihood(obs)
    print('c3 ok', float(r['loglik'].sum()))
    if len(sys.argv) > 2 and sys.argv[2] == 'down':
        r = mjp.expected_history_statistics(obs)
        print('c3 down ok', float(r['dwell'].sum()))
torch.cuda.synchronize()
