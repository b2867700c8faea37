import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from raoteh_b200 import engine, synth, tmjp
from raoteh_b200.lowering import TreeSchedule
dev = torch.device('cuda:0')
n_sites = int(os.environ.get('C5_SITES', 40000))
cfg = synth.config_c5(n_sites=n_sites)
sched = TreeSchedule(cfg['parent'], cfg['length'])
obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'], device=dev)
ch = tmjp.ToleranceChains(sched, cfg['Q'], cfg['pi'], dict(enumerate(cfg['part'])), cfg['rate_on'],
                          cfg['rate_off'], obs, n_chains=1, tol_obs=cfg['tol_obs'],
                          tol_obs_nodes=cfg['tol_obs_nodes'], cap_p=96, cap_t=48, seed=1, device=dev)
ch.initialize(); ch.sweep(5, stats=False); torch.cuda.synchronize()
for rep in range(3):
    ts = []
    for i in range(6):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(); ch._run(tmjp.MODE_SWEEP, n_sweeps=1, flags=3); ch.sweeps_done += 1
        b.record(); ch._run(tmjp.MODE_SUMMARY); c.record(); torch.cuda.synchronize()
        ts.append((round(a.elapsed_time(b), 2), round(b.elapsed_time(c), 2)))
    print(rep, ts)
