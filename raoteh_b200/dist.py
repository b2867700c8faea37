"""
Site / chain sharding across the GPUs of one box.

Sites are independent given (tree, Q, root prior) and chains are independent of
each other, so the path shards with NO data-path collective: rank r owns a
contiguous block of the site axis (or of the flattened (chain, site) axis) and
holds replicas of the tiny tree schedule, Q and per-edge P.  The only exchange
is one allreduce(sum, fp64) of [1 + S + S*S + S] values per evaluation
(sum log-lik, dwell[S], trans[S,S], root posterior sum[S]) over NCCL/NVLink.
Per-site outputs stay sharded.  Philox keys use GLOBAL trajectory indices, so
sampled histories do not depend on the number of ranks.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous block [lo, hi) of n_items owned by `rank` (sizes differ by at most 1)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def pack_stats(loglik_sum, dwell, trans, root_post_sum):
    """One flat fp64 vector [1 + S + S*S + S] for the allreduce."""
    return torch.cat([loglik_sum.reshape(1), dwell.reshape(-1), trans.reshape(-1),
                      root_post_sum.reshape(-1)])


def unpack_stats(vec, S):
    o = 0
    out = {}
    out['loglik_sum'] = vec[o]
    o += 1
    out['dwell'] = vec[o:o + S]
    o += S
    out['trans'] = vec[o:o + S * S].reshape(S, S)
    o += S * S
    out['root_post_sum'] = vec[o:o + S]
    return out


def allreduce_stats(vec):
    """In-place sum over ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


class ShardedEvaluation(object):
    """Site-sharded log-likelihood + expected history statistics.

    `evaluate_local(lo, hi)` must return a dict with 'loglik' [hi-lo] and the local
    sums 'dwell', 'trans', 'root_post_sum' as tensors; this class adds the shard
    bookkeeping and the single allreduce.  With raoteh_b200.engine:

        se = ShardedEvaluation(n_sites, S, lambda lo, hi: mjp.expected_history_statistics(obs_shard))
    """

    def __init__(self, n_sites, S, evaluate_local):
        self.n_sites = int(n_sites)
        self.S = int(S)
        self.evaluate_local = evaluate_local
        self.rank, self.world_size = world()
        self.lo, self.hi = shard_range(self.n_sites, self.rank, self.world_size)

    def __call__(self):
        r = self.evaluate_local(self.lo, self.hi)
        vec = pack_stats(r['loglik'].sum(), r['dwell'], r['trans'], r['root_post_sum'])
        allreduce_stats(vec)
        out = unpack_stats(vec, self.S)
        out['loglik_local'] = r['loglik']
        out['site_range'] = (self.lo, self.hi)
        return out


# ---------------------------------------------------------------------------------------
# sharded samplers (Rao-Teh chains, tolerance chains)
# ---------------------------------------------------------------------------------------
def shard_trajectories(n_chains, n_sites, rank=None, world_size=None):
    """Contiguous block (traj0, n_traj) of the flattened (chain, site) axis for `rank`.
    Pass it to RaoTehChains / ToleranceChains as traj0= and n_traj=; the Philox key of a
    trajectory is its GLOBAL index, so the sampled histories are those of one big run."""
    lo, hi = shard_range(int(n_chains) * int(n_sites), rank, world_size)
    return lo, hi - lo


def pack_sampler_stats(chains):
    """One flat fp64 vector of every accumulated sufficient statistic of a sampler shard:
    RaoTehChains -> [dwell S | trans S*S]; ToleranceChains -> [prim_dwell S | prim_trans S*S |
    tol_stats 4*n_parts | summary_sum 8]."""
    if hasattr(chains, 'prim_dwell'):
        parts = [chains.prim_dwell, chains.prim_trans, chains.tol_stats, chains.summary_sum]
    else:
        parts = [chains.dwell_sum, chains.trans_sum]
    return torch.cat([p.reshape(-1) for p in parts])


def unpack_sampler_stats(vec, S, n_parts=None):
    o = 0
    out = {}
    out['dwell'] = vec[o:o + S]
    o += S
    out['trans'] = vec[o:o + S * S].reshape(S, S)
    o += S * S
    if n_parts is not None:
        out['tol_stats'] = vec[o:o + 4 * n_parts].reshape(n_parts, 4)
        o += 4 * n_parts
        out['summary_sum'] = vec[o:o + 8]
    return out


def allreduce_sampler_stats(chains):
    """The sampler path's only collective: one allreduce(sum, fp64) of the packed statistics."""
    vec = pack_sampler_stats(chains)
    allreduce_stats(vec)
    return unpack_sampler_stats(vec, chains.S, getattr(chains, 'n_parts', None))
