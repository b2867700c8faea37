"""Benchmark legs behind bench.py `extra`: C4 (Rao-Teh sweeps, BASELINE.json configs[3]), C5
(tolerance model: likelihood, blocked Gibbs sampler, summary) and 61-state Rao-Teh.  Lives
beside bench.py, not in the package: its CPU legs run the oracle port."""
from __future__ import annotations

import time

import numpy as np
import torch


def _sync_max_ms(ms, dev, world):
    """max over ranks of a device-timed duration"""
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def _barrier(world):
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def ncu_counter(kernel_substr, key, default=None):
    """one counter of a kernel from the committed ncu summaries (newest round first)"""
    import json
    import os
    root = os.path.dirname(os.path.abspath(__file__))
    for name in ('r2_ncu_full_summary.json', 'r1_ncu_full_summary.json'):
        path = os.path.join(root, 'profiles', name)
        if not os.path.exists(path):
            continue
        for k in json.load(open(path)):
            if kernel_substr in k['kernel'] and key in k:
                try:
                    return float(str(k[key]).split()[0]), name
                except ValueError:
                    continue
    return default, None


def bench_c4_sharded(dev, rank, world, n_chains=4096, n_sites=10_000, burn_in=10, timed_sweeps=100,
                     cap=112, cpu_leg=True, time_dtype='float32'):
    """BASELINE.json configs[3] at its configured size: 4096 chains x 1e4 sites = 4.096e7
    (chain, site) trajectories of the 4-state HKY MJP on the 64-leaf tree, sharded over the ranks
    by dist.shard_trajectories (STRONG scaling: the total is fixed), `timed_sweeps` Rao-Teh sweeps
    of every trajectory in the timed region, then the sampler path's one allreduce of the
    sufficient statistics.  Reference loop: raoteh/sampler/_sampler.py:300-390."""
    from raoteh_b200 import dist as rdist
    from raoteh_b200 import engine, synth
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.raoteh import RaoTehChains
    cfg = synth.config_c4(n_sites=n_sites)
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    codes_pinned = torch.from_numpy(cfg['codes']).pin_memory()
    traj0, n_traj = rdist.shard_trajectories(n_chains, n_sites, rank, world)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # e2e: the leaf codes cross PCIe inside the timed region (they are per SITE, shared by chains)
    _barrier(world)
    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
    e0.record()
    obs = engine.Observations.from_leaf_codes(sched, codes_pinned.to(dev, non_blocking=True), cfg['leaves'],
                                              device=dev)
    e1.record()
    ch = RaoTehChains(sched, cfg['Q'], obs, n_chains=n_chains, root_distn=cfg['pi'], seed=20260204,
                      cap=cap, device=dev, traj0=traj0, n_traj=n_traj, time_dtype=time_dtype)
    k = ch.initialize()
    ch.sweep(burn_in, stats=False, auto_grow=True)
    _barrier(world)
    a, b = ev(), ev()
    a.record()
    ch.sweep(timed_sweeps, stats=True, auto_grow=True)
    b.record()
    _barrier(world)
    ms = _sync_max_ms(a.elapsed_time(b), dev, world)
    e2.record()
    red = rdist.allreduce_sampler_stats(ch)        # the sampler path's only collective
    host_stats = torch.cat([red['dwell'].reshape(-1), red['trans'].reshape(-1)]).cpu()
    e3.record()
    torch.cuda.synchronize()
    ms_e2e = _sync_max_ms(e0.elapsed_time(e1) + a.elapsed_time(b) + e2.elapsed_time(e3), dev, world)
    total = n_chains * n_sites
    sweeps = float(total) * timed_sweeps
    jumps = torch.tensor([float(ch.ev_total.double().sum())], dtype=torch.float64, device=dev)
    rdist.allreduce_stats(jumps)
    mean_jumps = float(jumps[0]) / total
    n = sched.n
    tb = 8 if time_dtype == 'float64' else 4
    # trajectory state streamed per sweep: node states + per-edge jump counts +
    # the jump list (time, parent-side state) + total, read and written once
    state_bytes = 2 * n + 4 + mean_jumps * (tb + 1)
    alloc_bytes = 2 * n + 4 + 4 + cap * (tb + 1) + 1
    lanes, src = ncu_counter('raoteh_kernel<4', 'smsp__thread_inst_executed_per_inst_executed.ratio')
    issue, _ = ncu_counter('raoteh_kernel<4', 'smsp__issue_active.avg.pct_of_peak_sustained_active')
    out = dict(
        workload='C4 (BASELINE configs[3]): Rao-Teh sweeps, 4-state HKY, 64-leaf tree, %d chains x %d '
                 'sites = %d trajectories in total, %d timed sweeps each after %d burn-in'
                 % (n_chains, n_sites, total, timed_sweeps, burn_in),
        scaling='strong', n_gpus=world, trajectories_total=total, trajectories_this_rank=n_traj,
        value=sweeps / (ms * 1e-3), unit='sweeps/s', ms=ms, timed_sweeps=timed_sweeps,
        time_dtype=time_dtype, cap=ch.cap, init_events_per_edge=k,
        state_bytes_per_trajectory_allocated=alloc_bytes,
        state_gb_this_rank=alloc_bytes * n_traj / 1e9,
        mean_real_jumps_per_trajectory=mean_jumps,
        expected_candidate_events_per_sweep=float(ch.omega * cfg['length'].sum()) + mean_jumps,
        hbm_gbs=sweeps * 2 * state_bytes / (ms * 1e-3) / 1e9,
        hbm_bytes_per_sweep=2 * state_bytes,
        lanes_active=lanes, issue_active=issue, ncu_source=src,
        e2e=dict(value=sweeps / (ms_e2e * 1e-3), unit='sweeps/s', ms=ms_e2e,
                 h2d_bytes_per_step=int(codes_pinned.numel()),
                 d2h_bytes_per_step=int(host_stats.numel() * 8)),
        reduced_dwell_sum=[float(x) for x in red['dwell'].cpu()],
        reduced_trans_sum=[float(x) for x in red['trans'].reshape(-1).cpu()],
        dwell_per_sweep_per_site_minus_tree_length=float(red['dwell'].sum()) / sweeps - float(cfg['length'].sum()))
    del ch, obs
    torch.cuda.empty_cache()
    if cpu_leg and rank == 0 and world == 1:
        out['cpu_port'] = _cpu_c4(cfg, sched)
    return out


def _cpu_c4(cfg, sched, n=30):
    """numpy restatement of the sweep (oracle/np_oracle.py), one trajectory, one core"""
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
        for _ in range(n):
            traj = np_oracle.raoteh_sweep(cfg['parent'], cfg['length'], B, rates, allowed,
                                          cfg['pi'], traj, rng)
        dt = (time.perf_counter() - t0) / n
        return dict(value=1.0 / dt, unit='sweeps/s', cores=1, kind='port',
                    sample='%d sweeps of one C4 trajectory, oracle/np_oracle.py raoteh_sweep' % n)
    except Exception as e:   # pragma: no cover
        return dict(error=repr(e))


def bench_c5_sharded(dev, rank, world, n_sites=1_000_000, burn_in=3, timed_sweeps=20, cpu_leg=True):
    """BASELINE.json configs[4]: p53-style tolerance MJP (61 codons x 20 tolerance classes), 1e6
    sites in total sharded over the ranks (STRONG scaling): (a) pruning log-likelihood under the
    primary proposal model + allreduce of the sum, (b) blocked Gibbs sweeps of the compound process
    (raoteh/sampler/_sample_tmjp_dense.py:40-171) + allreduce of the statistics."""
    from raoteh_b200 import dist as rdist
    from raoteh_b200 import engine, synth
    from raoteh_b200.lowering import TreeSchedule
    from raoteh_b200.tmjp import ToleranceChains
    cfg = synth.config_c5(n_sites=n_sites)
    sched = TreeSchedule(cfg['parent'], cfg['length'])
    ev = lambda: torch.cuda.Event(enable_timing=True)
    lo, hi = rdist.shard_range(n_sites, rank, world)
    out = dict(workload='C5 (BASELINE configs[4]): 61 codons x 20 tolerance classes, 25-taxon tree, '
                        '%d sites in total' % n_sites, scaling='strong', n_gpus=world,
               sites_this_rank=hi - lo)
    # (a) likelihood of the rank's site shard
    obs_ll = engine.Observations.from_leaf_codes(sched, np.ascontiguousarray(cfg['codes'][:, lo:hi]),
                                                 cfg['leaves'], device=dev)
    mjp = engine.TreeMJP(sched, cfg['Q_proposal'], root_distn=cfg['pi'], device=dev)
    ll = torch.empty(hi - lo, dtype=torch.float64, device=dev)
    st = torch.empty(hi - lo, dtype=torch.int8, device=dev)
    llsum = torch.zeros(1, dtype=torch.float64, device=dev)
    mjp.transition_matrices()
    for _ in range(2):
        mjp.log_likelihood(obs_ll, out=(ll, st))
    ts = []
    for _ in range(3):
        _barrier(world)
        llsum.zero_()
        a, b = ev(), ev()
        a.record()
        mjp.log_likelihood(obs_ll, out=(ll, st), loglik_sum=llsum)
        rdist.allreduce_stats(llsum)
        b.record()
        _barrier(world)
        ts.append(_sync_max_ms(a.elapsed_time(b), dev, world))
    ms = float(np.mean(ts))
    out['loglik'] = dict(value=n_sites * sched.n_edges / (ms * 1e-3), unit='messages/s', ms=ms,
                         loglik_sum=float(llsum[0]), finite=bool(torch.isfinite(ll).all()))
    del obs_ll, mjp, ll, st
    # (b) blocked Gibbs sampler: trajectories = sites (one chain), global site index = trajectory index
    obs = engine.Observations.from_leaf_codes(sched, cfg['codes'], cfg['leaves'], device=dev)
    ch = ToleranceChains(sched, cfg['Q'], cfg['pi'], dict(enumerate(cfg['part'])), cfg['rate_on'],
                         cfg['rate_off'], obs, n_chains=1, tol_obs=cfg['tol_obs'],
                         tol_obs_nodes=cfg['tol_obs_nodes'], cap_p=96, cap_t=48, seed=20260205,
                         device=dev, traj0=lo, n_traj=hi - lo)
    k = ch.initialize()
    ch.sweep(burn_in, stats=False)
    _barrier(world)
    a, b = ev(), ev()
    a.record()
    ch.sweep(timed_sweeps, stats=True)
    b.record()
    _barrier(world)
    ms = _sync_max_ms(a.elapsed_time(b), dev, world)
    red = rdist.allreduce_sampler_stats(ch)
    sweeps = float(n_sites) * timed_sweeps
    issue, src = ncu_counter('tmjp_kernel', 'smsp__issue_active.avg.pct_of_peak_sustained_active')
    out['gibbs'] = dict(value=sweeps / (ms * 1e-3), unit='sweeps/s', ms=ms, timed_sweeps=timed_sweeps,
                        time_dtype='float32', init_events_per_edge=k, issue_active=issue, ncu_source=src,
                        mean_primary_jumps=float(ch.p_total.double().mean()),
                        reduced_primary_dwell_total=float(red['dwell'].sum()),
                        reduced_primary_transitions_total=float(red['trans'].sum()),
                        reduced_tolerance_stats_total=[float(x) for x in red['tol_stats'].sum(dim=0).cpu()])
    # Rao-Blackwellised summary of every current trajectory
    ch.tolerance_summary()
    _barrier(world)
    a, b = ev(), ev()
    a.record()
    ch.tolerance_summary()
    b.record()
    _barrier(world)
    ms = _sync_max_ms(a.elapsed_time(b), dev, world)
    out['summary'] = dict(value=n_sites / (ms * 1e-3), unit='trajectories/s', ms=ms)
    del ch, obs
    torch.cuda.empty_cache()
    if cpu_leg and rank == 0 and world == 1:
        try:
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
            n_cpu = 40
            for _ in range(n_cpu):
                prim, tols = np_tmjp.gibbs_sweep(*margs, prim, tols, rng)
            out['gibbs']['cpu_port'] = dict(value=n_cpu / (time.perf_counter() - t0), unit='sweeps/s', cores=1,
                                            kind='port', sample='%d sweeps of one C5 trajectory, '
                                            'oracle/np_tmjp.py gibbs_sweep' % n_cpu)
        except Exception as e:   # pragma: no cover
            out['gibbs']['cpu_port'] = dict(error=repr(e))
    return out


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
