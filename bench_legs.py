"""Benchmark legs behind bench.py `extra`: C4 (Rao-Teh sweeps, BASELINE.json configs[3]), C5
(tolerance model: likelihood, blocked Gibbs sampler, summary) and 61-state Rao-Teh.  Lives
beside bench.py, not in the package: its CPU legs run the oracle port."""
from __future__ import annotations

import time

import numpy as np
import torch


def bench_c4(dev, args, n_chains=128, n_sites=10_000, sweeps_per_launch=25, launches=4):
    from raoteh_b200 import engine, synth
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
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


def bench_c5(dev, args, n_sites=100_000, sweeps_per_launch=5, launches=3, loglik_sites=1_000_000):
    """C5 legs: (a) 61-state pruning log-likelihood under the primary proposal model on the
    25-taxon tree, (b) blocked Gibbs sweeps of the compound tolerance process (61 codons x 20
    amino-acid classes) with the Rao-Blackwellised tolerance summary fused after every sweep."""
    from raoteh_b200 import engine, synth
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.tmjp import ToleranceChains
    cfg = synth.config_c5(n_sites=max(n_sites, loglik_sites))
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    out = dict(workload='C5: 61 codons x 20 tolerance classes, 25-taxon tree (48 edges x 0.1)')
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # (a) likelihood
    obs_ll = engine.Observations.from_leaf_codes(sched, cfg['codes'][:, :loglik_sites], cfg['leaves'], device=dev)
    mjp = engine.TreeMJP(sched, cfg['Q_proposal'], root_distn=cfg['pi'], device=dev)
    ll = torch.empty(loglik_sites, dtype=torch.float64, device=dev)
    st = torch.empty(loglik_sites, dtype=torch.int8, device=dev)
    mjp.transition_matrices()
    for _ in range(2):
        mjp.log_likelihood(obs_ll, out=(ll, st))
    ts = []
    for _ in range(3):
        a, b = ev(), ev()
        a.record()
        mjp.log_likelihood(obs_ll, out=(ll, st))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    out['loglik'] = dict(n_sites=loglik_sites, ms=ms, messages_per_sec=loglik_sites * sched.n_edges / (ms * 1e-3),
                         finite=bool(torch.isfinite(ll).all()))
    del obs_ll, mjp, ll, st
    # (b) blocked Gibbs sampler
    obs = engine.Observations.from_leaf_codes(sched, cfg['codes'][:, :n_sites], cfg['leaves'], device=dev)
    ch = ToleranceChains(sched, cfg['Q'], cfg['pi'], dict(enumerate(cfg['part'])), cfg['rate_on'],
                         cfg['rate_off'], obs, n_chains=1, tol_obs=cfg['tol_obs'][:, :, :n_sites],
                         tol_obs_nodes=cfg['tol_obs_nodes'], cap_p=96, cap_t=48, seed=20260205, device=dev)
    k = ch.initialize()
    ch.sweep(3, stats=False)
    ch.sweep(1, stats=False, summary=True)    # warm-up of the summary path (scratch pool growth)
    torch.cuda.synchronize()
    res = {}
    for name, summary in (('sweep', False), ('sweep_plus_summary', True)):
        ts = []
        for _ in range(launches):
            a, b = ev(), ev()
            a.record()
            ch.sweep(sweeps_per_launch, stats=True, summary=summary)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.mean(ts))
        res[name] = dict(ms_per_launch=ms, sweeps_per_sec=ch.n_traj * sweeps_per_launch / (ms * 1e-3))
    a, b = ev(), ev()
    a.record()
    ch.tolerance_summary()
    b.record()
    torch.cuda.synchronize()
    res['summary_only'] = dict(ms=a.elapsed_time(b), trajectories_per_sec=ch.n_traj / (a.elapsed_time(b) * 1e-3))
    # CPU leg: pure-Python/numpy restatement of the same sweep, one trajectory, one core
    try:
        import time
        from oracle import np_tmjp
        rng = np.random.default_rng(0)
        nodes = dict((int(v), int(cfg['codes'][i, 0])) for i, v in enumerate(cfg['leaves']))
        dd = [dict() for _ in range(cfg['n_parts'])]
        for c in range(cfg['n_parts']):
            bits = int(cfg['tol_obs'][0, c, 0])
            dd[c][int(cfg['tol_obs_nodes'][0])] = set(s for s in (0, 1) if (bits >> s) & 1)
        margs = (cfg['parent'], cfg['length'], cfg['Q'], cfg['part'], cfg['n_parts'], cfg['pi'],
                 cfg['rate_on'], cfg['rate_off'], nodes, dd)
        prim, tols = np_tmjp.gibbs_init(*margs, rng)
        t0 = time.perf_counter()
        n_cpu = 60
        for _ in range(n_cpu):
            prim, tols = np_tmjp.gibbs_sweep(*margs, prim, tols, rng)
        res['cpu_port_sweeps_per_sec_1core'] = n_cpu / (time.perf_counter() - t0)
        res['cpu_port_sample'] = '%d sweeps of one C5 trajectory, oracle/np_tmjp.py gibbs_sweep' % n_cpu
    except Exception as e:   # pragma: no cover
        res['cpu_port_error'] = repr(e)
    out['gibbs'] = dict(n_trajectories=ch.n_traj, init_events_per_edge=k, sweeps_per_launch=sweeps_per_launch,
                        mean_primary_jumps=float(ch.p_total.double().mean()),
                        mean_tolerance_toggles_per_class=float(ch.t_total.double().mean()),
                        omega_primary=ch.omega_p, omega_tolerance=ch.omega_t, **res)
    return out


def bench_codon_raoteh(dev, args, n_sites=20_000, n_chains=4, sweeps_per_launch=5, launches=3):
    """Plain Rao-Teh sweeps of the 61-state codon model on the C3 tree (warp-per-trajectory)."""
    from raoteh_b200 import engine, synth
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
    cfg = synth.config_c3(n_sites=n_sites)
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'], device=dev)
    omega = 2 * float(np.max(-np.diag(cfg['Q'])))
    cap = int(max(4 * omega * cfg["length"].sum() + 64, 2 * (sched.n - 1)))   # room for the initial history
    ch = RaoTehChains(sched, cfg['Q'], obs, n_chains=n_chains, root_distn=cfg['pi'], seed=20260202,
                      cap=cap, device=dev)
    k = ch.initialize()
    ch.sweep(3, stats=False)
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
    return dict(workload='61-state codon Rao-Teh, 128-leaf tree, %d chains x %d sites' % (n_chains, n_sites),
                n_trajectories=ch.n_traj, cap=cap, init_events_per_edge=k, ms_per_launch=ms,
                sweeps_per_sec=ch.n_traj * sweeps_per_launch / (ms * 1e-3),
                mean_real_jumps_per_trajectory=float(ch.ev_total.double().mean()),
                expected_candidate_events_per_sweep=float(omega * cfg['length'].sum()))


def bench_next_rows(dev, args):
    """Measurements for the rows SURVEY.md 8(f) marks 'next': per-site per-branch expectations
    (f2) at C2 and C3 size, and the device-side Metropolis-Hastings pipeline (f1) at C5 size."""
    from raoteh_b200 import engine, synth
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.mh import ToleranceMetropolisChains
    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = {}
    for name, cfg in (('c2', synth.config_c2(n_sites=1_000_000)), ('c3', synth.config_c3(n_sites=100_000))):
        sched = TreeSchedule(cfg['parent'], cfg['length'])
        mjp = engine.TreeMJP(sched, cfg['Q'], root_distn=cfg['pi'], device=dev)
        obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'], device=dev)
        K = mjp.transition_kernels(None)
        for _ in range(2):
            r = mjp.branch_expectations(obs, K=K)
        ts = []
        for _ in range(3):
            a, b = ev(), ev()
            a.record()
            r = mjp.branch_expectations(obs, K=K)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.mean(ts))
        out['branch_expectations_' + name] = dict(
            workload='%s: up pass + down pass with one expected count per site and branch '
                     '([n_nodes, n_sites] fp64 out)' % name.upper(),
            ms=ms, messages_per_sec=obs.n_sites * sched.n_edges / (ms * 1e-3),
            out_bytes=int(r['branch'].numel() * 8))
        del mjp, obs, r
    cfg = synth.config_c5(n_sites=20_000)
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'], device=dev)
    mh = ToleranceMetropolisChains(sched, cfg['Q'], cfg['pi'], dict(enumerate(int(p) for p in cfg['part'])),
                                   cfg['rate_on'], cfg['rate_off'], obs, n_chains=1, cap=192,
                                   seed=20260205, device=dev)
    mh.step(3, stats=False)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    mh.step(5)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out['mh_c5'] = dict(workload='C5: Rao-Teh proposal under the approximate primary process + both '
                                 'log-likelihoods + tolerance summary + accept/reject, 2e4 trajectories',
                        ms_per_step=ms, steps_per_sec=mh.n_traj / (ms * 1e-3),
                        acceptance=mh.n_accepted / max(1, mh.n_proposed))
    return out
