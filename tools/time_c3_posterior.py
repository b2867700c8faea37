import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule
dev = torch.device('cuda:0')
cfg = synth.config_c3(n_sites=100_000)
sched = TreeSchedule(cfg['parent'], cfg['length'])
mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'], device=dev)
obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'], device=dev)
mjp.events = {}
for _ in range(2):
    r = mjp.expected_history_statistics(obs)
torch.cuda.synchronize()
mjp.events = sub = {}
ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = mjp.expected_history_statistics(obs); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
pair = lambda L: [(L[i], L[i + 1]) for i in range(0, len(L) - 1, 2)]
up = np.mean([x.elapsed_time(y) for x, y in pair(sub['up'])])
down = np.mean([x.elapsed_time(y) for x, y in pair(sub['down'])])
n_int = int((~sched.is_leaf[1:]).sum())
print(json.dumps(dict(total_ms=float(np.mean(ts)), up_ms=float(up), down_ms=float(down), n_levels=r['n_levels'],
                      edges=sched.n_edges, internal_edges=n_int,
                      down_tflops_executed=float(obs.n_sites * (3 * n_int + (sched.n_edges - n_int)) * 2 * 64 * 64 / (down * 1e-3) / 1e12))))
