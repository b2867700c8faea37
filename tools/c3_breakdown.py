"""C3 expectations: time per stage (CUDA events around each library call)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule
cfg = synth.config_c3(n_sites=100_000)
sched = TreeSchedule(cfg['parent'], cfg['length'])
mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'])
obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'])
ev = lambda: torch.cuda.Event(enable_timing=True)
for _ in range(2):
    mjp.expected_history_statistics(obs)
torch.cuda.synchronize()
res = {}
for rep in range(3):
    mjp._P_valid = False
    marks = [ev() for _ in range(5)]
    marks[0].record()
    mjp.transition_matrices()
    marks[1].record()
    up = mjp.log_likelihood(obs, keep_partials=True)
    marks[2].record()
    post = mjp.posterior(obs, want_node_distn=False)      # includes another up pass
    marks[3].record()
    mjp.history_statistics(post['W'])
    marks[4].record()
    torch.cuda.synchronize()
    for name, a, b in (('expm', 0, 1), ('up_store', 1, 2), ('up+down', 2, 3), ('frechet+accumulate', 3, 4)):
        res.setdefault(name, []).append(marks[a].elapsed_time(marks[b]))
for k, v in res.items():
    print('%-20s %.3f ms' % (k, np.mean(v)))
