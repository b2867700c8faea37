"""C4 benchmark leg: Rao-Teh Gibbs sweeps, 4-state HKY on a 64-leaf tree
(BASELINE.json configs[3]); used by bench.py `extra`."""
from __future__ import annotations

import time

import numpy as np
import torch


def bench_c4(dev, args, n_chains=128, n_sites=10_000, sweeps_per_launch=25, launches=4):
    from . import engine, synth
    from .lowering import TreeSchedule
    from .raoteh import RaoTehChains
    cfg = synth.config_c4(n_sites=n_sites)
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'], device=dev)
    ch = RaoTehChains(sched, cfg['Q'], obs, n_chains=n_chains, root_distn=cfg['pi'], seed=20260204,
                      cap=96, device=dev)
    k = ch.initialize()
    ch.sweep(10, stats=False)          # burn-in / warm-up
    torch.cuda.synchronize()
    ts = []
    for _ in range(launches):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ch.sweep(sweeps_per_launch)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ch.check()
    ms = float(np.mean(ts))
    n_traj = ch.n_traj
    sweeps = n_traj * sweeps_per_launch
    mean_events = float(ch.ev_total.double().mean())
    out = dict(workload='C4: Rao-Teh sweeps, 4-state HKY, 64-leaf tree, %d chains x %d sites '
                        '(timed window of %d sweeps per trajectory per launch)'
                        % (n_chains, n_sites, sweeps_per_launch),
               n_trajectories=n_traj, ms_per_launch=ms, sweeps_per_sec=sweeps / (ms * 1e-3),
               mean_real_jumps_per_trajectory=mean_events, init_events_per_edge=k,
               expected_candidate_events_per_sweep=float(ch.omega * cfg['length'].sum()))
    # CPU leg: numpy restatement of the sweep, one trajectory, single core
    try:
        from oracle import np_oracle
        rng = np.random.default_rng(0)
        S = 4
        omega, B, rates = np_oracle.uniformized(cfg['Q'], 2.0)
        allowed = np.ones((sched.n, S))
        for i, v in enumerate(cfg['leaves']):
            allowed[v] = 0
            allowed[v, cfg['codes'][i, 0]] = 1
        traj = np_oracle.raoteh_init(cfg['parent'], cfg['length'], B, allowed, cfg['pi'], rng)
        t0 = time.perf_counter()
        n = 30
        for _ in range(n):
            traj = np_oracle.raoteh_sweep(cfg['parent'], cfg['length'], B, rates, allowed,
                                          cfg['pi'], traj, rng)
        dt = (time.perf_counter() - t0) / n
        out['cpu_port_sweeps_per_sec_1core'] = 1.0 / dt
    except Exception as e:   # pragma: no cover
        out['cpu_port_error'] = repr(e)
    return out


if __name__ == '__main__':
    import json
    import sys
    print(json.dumps(bench_c4(torch.device('cuda:0'), None,
                              n_chains=int(sys.argv[1]) if len(sys.argv) > 1 else 128)))
