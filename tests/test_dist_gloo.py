"""N > 1 host logic on CPU: world_size-2 gloo run of the site-sharded evaluation
(shard ranges, packed allreduce of [1 + S + S*S + S] doubles).  The per-shard
numbers come from the oracle so that no GPU is needed; the sharded result must
equal the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raoteh_b200 import dist as rdist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100, 1_000_003):
        for w in (1, 2, 3, 8):
            blocks = [rdist.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    S = 4
    ll = torch.tensor(-12.5, dtype=torch.float64)
    dwell = torch.arange(S, dtype=torch.float64)
    trans = torch.arange(S * S, dtype=torch.float64).reshape(S, S)
    rp = torch.ones(S, dtype=torch.float64)
    out = rdist.unpack_stats(rdist.pack_stats(ll, dwell, trans, rp), S)
    assert float(out['loglik_sum']) == -12.5
    assert torch.equal(out['dwell'], dwell) and torch.equal(out['trans'], trans)
    assert torch.equal(out['root_post_sum'], rp)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, n_sites, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    from oracle import np_oracle
    from raoteh_b200 import synth
    cfg = synth.config_c2(n_sites=n_sites, n_leaves=8)
    P = np_oracle.expm_edges(cfg['Q'], cfg['length'])

    def evaluate_local(lo, hi):
        obs = np_oracle.Obs('codes', 4, hi - lo, leaf_nodes=cfg['leaves'], codes=cfg['codes'][:, lo:hi])
        r = np_oracle.expected_history_statistics(cfg['parent'], cfg['length'], cfg['Q'], P, obs, cfg['pi'])
        return dict(loglik=torch.from_numpy(r['loglik']), dwell=torch.from_numpy(r['dwell']),
                    trans=torch.from_numpy(r['trans']),
                    root_post_sum=torch.from_numpy(r['root_post'].sum(axis=0)))
    se = rdist.ShardedEvaluation(n_sites, 4, evaluate_local)
    out = se()
    q.put((rank, se.lo, se.hi, float(out['loglik_sum']), out['dwell'].numpy().copy(),
           out['trans'].numpy().copy(), out['root_post_sum'].numpy().copy()))
    dist.destroy_process_group()


def test_sharded_evaluation_two_ranks_gloo():
    from oracle import np_oracle
    from raoteh_b200 import synth
    n_sites = 301
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_sites, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 151, 151, 301)
    cfg = synth.config_c2(n_sites=n_sites, n_leaves=8)
    P = np_oracle.expm_edges(cfg['Q'], cfg['length'])
    obs = np_oracle.Obs('codes', 4, n_sites, leaf_nodes=cfg['leaves'], codes=cfg['codes'])
    full = np_oracle.expected_history_statistics(cfg['parent'], cfg['length'], cfg['Q'], P, obs, cfg['pi'])
    for r in res:      # both ranks hold the same reduced statistics
        np.testing.assert_allclose(r[3], full['loglik'].sum(), rtol=1e-12)
        np.testing.assert_allclose(r[4], full['dwell'], rtol=1e-12)
        np.testing.assert_allclose(r[5], full['trans'], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(r[6], full['root_post'].sum(axis=0), rtol=1e-12)


class _FakeChains(object):
    """Host stand-in with the statistic tensors of raoteh_b200.tmjp.ToleranceChains."""

    def __init__(self, S, n_parts, traj0, n_traj):
        g = torch.Generator().manual_seed(0)
        # statistics are sums over trajectories: give trajectory t the contribution f(t)
        t = torch.arange(traj0, traj0 + n_traj, dtype=torch.float64)
        self.S, self.n_parts = S, n_parts
        self.prim_dwell = torch.stack([(t * (s + 1)).sum() for s in range(S)])
        self.prim_trans = torch.stack([(t % (k + 2)).sum() for k in range(S * S)]).reshape(S, S)
        self.tol_stats = torch.stack([(t + k).sum() for k in range(4 * n_parts)]).reshape(n_parts, 4)
        self.summary_sum = torch.stack([(t * 0 + 1).sum() * (k + 1) for k in range(8)])


def _sampler_worker(rank, world_size, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    traj0, n_traj = rdist.shard_trajectories(3, 101)
    out = rdist.allreduce_sampler_stats(_FakeChains(5, 3, traj0, n_traj))
    q.put((rank, traj0, n_traj, out['dwell'].numpy().copy(), out['trans'].numpy().copy(),
           out['tol_stats'].numpy().copy(), out['summary_sum'].numpy().copy()))
    dist.destroy_process_group()


def test_sharded_sampler_statistics_two_ranks_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampler_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 152, 152, 151)
    full = _FakeChains(5, 3, 0, 303)
    for r in res:
        np.testing.assert_allclose(r[3], full.prim_dwell.numpy(), rtol=1e-12)
        np.testing.assert_allclose(r[4], full.prim_trans.numpy(), rtol=1e-12)
        np.testing.assert_allclose(r[5], full.tol_stats.numpy(), rtol=1e-12)
        np.testing.assert_allclose(r[6], full.summary_sum.numpy(), rtol=1e-12)
