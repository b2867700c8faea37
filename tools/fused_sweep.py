"""Timing of the fused S <= 4 evaluation against the two-kernel path at C2 size, for a few
CTAs-per-SM settings (the L2 footprint of the per-CTA scratch is grid x 254 KB)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule

cfg = synth.config_c2(n_sites=1_000_000)
sched = TreeSchedule(cfg['parent'], cfg['length'])
obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'])
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timeit(mjp, n=10):
    for _ in range(3):
        mjp.posterior(obs, want_node_distn=False)
    ts = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = mjp.posterior(obs, want_node_distn=False)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts)), r


mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'])
mjp.fused = False
m, mn, r0 = timeit(mjp)
print('two kernels: mean %.3f min %.3f ms' % (m, mn), flush=True)
W0 = r0['W'].clone(); ll0 = r0['loglik'].clone()
for ctas in (0, 4, 3, 2, 1):
    mjp.fused = True
    mjp.fused_ctas_per_sm = ctas
    m, mn, r = timeit(mjp)
    dW = float((r['W'] - W0).abs().max() / W0.abs().max())
    dl = float((r['loglik'] - ll0).abs().max())
    print('fused ctas/sm %d: mean %.3f min %.3f ms   max rel dW %.2e  max d loglik %.2e' % (ctas, m, mn, dW, dl), flush=True)
