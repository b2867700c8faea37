"""Diagnostic: z-scores of the Rao-Teh sweep statistics against the closed form for growing trees."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from oracle import np_oracle
from raoteh_b200 import engine, synth
from raoteh_b200.lowering import TreeSchedule
from raoteh_b200.raoteh import RaoTehChains


def run(n_leaves, td, burn, n_sites=3, mean_len=0.1, groups=16, n_chains=512, n_sweeps=100, seed=20260204, uf=2.0):
    rng = np.random.default_rng(seed)
    parent, length, leaves = synth.random_binary_tree(n_leaves, mean_len, rng)
    Q, pi = synth.hky85()
    codes = synth.simulate_leaf_codes(parent, length, leaves, Q, pi, n_sites, rng, 0.0)
    sched = TreeSchedule(parent, length)
    obs = engine.Observations.from_leaf_codes(sched, codes, leaves)
    P = np_oracle.expm_edges(Q, length)
    o = np_oracle.expected_history_statistics(
        parent, length, Q, P, np_oracle.Obs('codes', 4, n_sites, leaf_nodes=leaves, codes=codes), pi)
    dwell = np.zeros((groups, 4)); trans = np.zeros((groups, 16))
    for g in range(groups):
        ch = RaoTehChains(sched, Q, obs, n_chains=n_chains, root_distn=pi, seed=5000 + g, cap=160, time_dtype=td,
                          uniformization_factor=uf)
        ch.sweep(burn, stats=False)
        ch.sweep(n_sweeps)
        dwell[g] = ch.dwell_sum.cpu().numpy() / (n_chains * n_sweeps)
        trans[g] = ch.trans_sum.cpu().numpy().reshape(-1) / (n_chains * n_sweeps)
    zs = []
    for got, want in ((dwell, o['dwell']), (trans, o['trans'].reshape(-1))):
        m = got.mean(axis=0); se = got.std(axis=0, ddof=1) / np.sqrt(groups)
        zs.append(np.where(se > 0, (m - want) / np.where(se > 0, se, 1), 0))
    print('leaves', n_leaves, td, 'burn', burn, 'uf', uf, 'minlen %.2e' % length[1:].min(),
          'z dwell', np.round(zs[0], 1), 'z trans max', np.round(np.abs(zs[1]).max(), 1),
          'rel dwell', np.round(dwell.mean(axis=0) / o['dwell'] - 1, 4), flush=True)


run(32, 'float32', 60, groups=32, n_sweeps=400)
run(32, 'float32', 1000, groups=32, n_sweeps=400)
run(32, 'float32', 3000, groups=32, n_sweeps=400)
run(32, 'float32', 1000, groups=32, n_sweeps=400, uf=4.0)
run(64, 'float32', 1000, groups=32, n_sweeps=400)
