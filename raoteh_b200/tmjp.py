"""
Batched tolerance-process sampler: many independent (chain, site) compound
trajectories on device, one warp per trajectory (csrc/rt_tmjp.cu).

Host side of SURVEY rows A17/A18.  Mirrors the generator loop of the reference,
raoteh/sampler/_sample_tmjp_dense.py:40-171 (gen_histories_v1): uniformization
constants of the primary and of the 2-state tolerance process (:78-96), initial
jointly feasible history (:72 -> get_feasible_history :509-627), then blocked
Gibbs sweeps (:116-171); per sampled history the statistics of
_mjp_dense.get_history_statistics (raoteh/sampler/_mjp_dense.py:150) and the
Rao-Blackwellised summary _tmjp_dense.get_tolerance_summary
(raoteh/sampler/_tmjp_dense.py:724-855), as examples/p53/blink.py:54-67 uses them.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native
from .engine import OBS_CODES, OBS_MASK, _ptr, _stream

MODE_INIT_PRIMARY, MODE_INIT_TOLERANCE, MODE_SWEEP, MODE_SUMMARY, MODE_TRAJ_LOGLIK = 0, 1, 2, 3, 4
F_STATS_PRIMARY, F_STATS_TOLERANCE, F_SUMMARY = 1, 2, 4
F_SKIP_PRIMARY, F_SKIP_TOLERANCE = 8, 16

SUMMARY_FIELDS = ('initial_on', 'initial_off', 'dwell_on', 'dwell_off',
                  'nabsorptions', 'ngains', 'nlosses')


def absorption_rates(Q_primary, primary_to_part, n_parts):
    """[S, n_parts]: sum of Q[s, s'] over s' != s in class c
    (raoteh/sampler/_tmjp_dense.py:929-962 for every class at once)."""
    Q = np.asarray(Q_primary, dtype=np.float64)
    S = Q.shape[0]
    part = np.asarray(primary_to_part)
    out = np.zeros((S, n_parts), dtype=np.float64)
    off = Q - np.diag(np.diag(Q))
    for c in range(n_parts):
        out[:, c] = off[:, part == c].sum(axis=1)
    return out


class ToleranceChains(object):
    """n_chains x n_sites compound (primary + n_parts tolerance) trajectories.

    trajectory t = chain * n_sites + site; primary observations and disease data
    are per site.  `tol_obs`: uint8 [n_tol_obs, n_parts, n_sites] with bit0 =
    'off allowed', bit1 = 'on allowed' for the nodes listed in `tol_obs_nodes`
    (the reference's disease_data: per class, node -> set of allowed states).
    """

    def __init__(self, sched, Q_primary, primary_distn, primary_to_part, rate_on, rate_off,
                 obs, n_chains=1, tol_obs=None, tol_obs_nodes=None, uniformization_factor=2.0,
                 cap_p=64, cap_t=32, seed=0, device='cuda', traj0=0, n_traj=None):
        if not torch.cuda.is_available():
            raise _native.NativeError('raoteh_b200 needs a CUDA device; there is no CPU fallback')
        if uniformization_factor <= 1:
            raise ValueError('the uniformization factor must be greater than 1')
        Q = np.asarray(Q_primary, dtype=np.float64)
        self.S = S = Q.shape[0]
        if not 2 <= S <= 64:
            raise ValueError('the number of primary states must be in 2..64')
        part = np.asarray([primary_to_part[s] for s in range(S)], dtype=np.int64)
        self.n_parts = NP = int(part.max()) + 1
        if NP > 32:
            raise ValueError('at most 32 tolerance classes')
        if obs.kind not in (OBS_CODES, OBS_MASK):
            raise ValueError('primary observations are hard codes or allowed-state masks')
        self.sched, self.obs = sched, obs
        self.device = dev = torch.device(device)
        self.n_sites = obs.n_sites
        self.n_chains = int(n_chains)
        self.traj0 = int(traj0)
        self.n_traj = T = int(n_chains) * self.n_sites if n_traj is None else int(n_traj)
        self.cap_p, self.cap_t, self.seed = int(cap_p), int(cap_t), int(seed)
        self.rate_on, self.rate_off = float(rate_on), float(rate_off)
        # uniformization (raoteh/sampler/_sample_tmjp_dense.py:78-96)
        q = -np.diag(Q)
        self.omega_p = float(uniformization_factor * q.max())
        self.omega_t = float(uniformization_factor * max(self.rate_on, self.rate_off))
        if not self.omega_p > 0:
            raise ValueError('the rate matrix is empty')
        self.Q_host = Q
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        B = np.eye(S) + Q / self.omega_p
        if primary_distn is not None:
            # the reference keeps a primary state with zero prior out of EVERY chunk, not only
            # the root's (raoteh/sampler/_sample_tmjp_dense.py:301-305): no uniformized step may
            # enter such a state
            B[:, np.asarray(primary_distn, dtype=np.float64) == 0] = 0.0
        self.B = to(B)
        self.rate_p = to(self.omega_p - q)
        self.pi_p = None if primary_distn is None else to(np.asarray(primary_distn, dtype=np.float64))
        self.part = to(part.astype(np.uint8))
        self.absorb = to(absorption_rates(Q, part, NP))
        ops, n_slots = sched.up_program(obs.obs_slot)
        self.ops, self.n_ops, self.n_slots = to(ops), len(ops), n_slots
        self.parent, self.length = to(sched.parent.copy()), to(sched.length.copy())
        self.tol_obs = self.tol_obs_slot = None
        if tol_obs is not None:
            slot = np.full(sched.n, -1, dtype=np.int32)
            slot[np.asarray(tol_obs_nodes)] = np.arange(len(tol_obs_nodes), dtype=np.int32)
            self.tol_obs = to(np.asarray(tol_obs, dtype=np.uint8))
            assert self.tol_obs.shape[1] == NP and self.tol_obs.shape[2] == self.n_sites
            self.tol_obs_slot = to(slot)
        n = sched.n
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        self.p_node, self.p_cnt = z((T, n), torch.uint8), z((T, n), torch.uint8)
        self.p_total = z(T, torch.int32)
        self.p_time, self.p_sb = z((T, self.cap_p), torch.float32), z((T, self.cap_p), torch.uint8)
        self.t_node = z((T, n), torch.int32)
        self.t_cnt = z((T, n, NP), torch.uint8)
        self.t_total = z((T, NP), torch.uint8)
        self.t_time = z((T, NP, self.cap_t), torch.float32)
        self.status = z(T, torch.int8)
        self.prim_dwell = z(S, torch.float64)
        self.prim_trans = z((S, S), torch.float64)
        self.tol_stats = z((NP, 4), torch.float64)
        self.summary_sum = z(8, torch.float64)
        self.summary_out = z((T, 8), torch.float64)
        self.traj_loglik = z(T, torch.float64)
        self.sweeps_done = 0
        self.initialized = False
        self._primary_src = None
        self.p_time64 = None      # fp64 jump times of caller-loaded trajectories (load_primary_trajectories)

    # -- the C-ABI call ------------------------------------------------------------
    def _args(self, mode, n_sweeps=1, init_k=0, flags=0):
        A = _native.TmjpArgs()
        A.S, A.n_parts, A.n_nodes = self.S, self.n_parts, self.sched.n
        A.n_ops, A.n_slots = self.n_ops, self.n_slots
        A.cap_p, A.cap_t, A.obs_kind = self.cap_p, self.cap_t, self.obs.kind
        A.program, A.parent, A.length = _ptr(self.ops), _ptr(self.parent), _ptr(self.length)
        A.B, A.rate_p, A.pi_p = _ptr(self.B), _ptr(self.rate_p), _ptr(self.pi_p)
        A.part, A.absorb = _ptr(self.part), _ptr(self.absorb)
        A.rate_on, A.rate_off, A.omega_t = self.rate_on, self.rate_off, self.omega_t
        A.omega_p = self.omega_p
        A.obs, A.obs_stride = _ptr(self.obs.data), self.obs.stride
        A.tol_obs, A.tol_obs_slot = _ptr(self.tol_obs), _ptr(self.tol_obs_slot)
        A.tol_obs_stride = 0 if self.tol_obs is None else int(self.tol_obs.shape[2])
        A.n_traj, A.n_sites, A.traj0 = self.n_traj, self.n_sites, self.traj0
        src = self._primary_src
        if src is None:
            A.p_node, A.p_cnt = _ptr(self.p_node), _ptr(self.p_cnt)
            A.pn_traj_stride, A.pn_node_stride = self.sched.n, 1
            A.p_total, A.p_time, A.p_sb = _ptr(self.p_total), _ptr(self.p_time), _ptr(self.p_sb)
        else:   # primary trajectories of a RaoTehChains (trajectory-minor node arrays)
            A.cap_p = src.cap
            A.p_node, A.p_cnt = _ptr(src.node_state), _ptr(src.ev_count)
            A.pn_traj_stride, A.pn_node_stride = 1, src.stride
            A.p_total, A.p_time, A.p_sb = _ptr(src.ev_total), _ptr(src.ev_time), _ptr(src.ev_sb)
        A.t_node, A.t_cnt = _ptr(self.t_node), _ptr(self.t_cnt)
        A.t_total, A.t_time = _ptr(self.t_total), _ptr(self.t_time)
        A.status = _ptr(self.status)
        A.seed, A.sweep0 = self.seed, self.sweeps_done
        A.n_sweeps, A.mode, A.init_k, A.flags = int(n_sweeps), mode, int(init_k), int(flags)
        A.prim_dwell, A.prim_trans = _ptr(self.prim_dwell), _ptr(self.prim_trans)
        A.tol_stats, A.summary_sum = _ptr(self.tol_stats), _ptr(self.summary_sum)
        A.summary_out = _ptr(self.summary_out)
        A.traj_loglik = _ptr(self.traj_loglik)
        A.p_time64 = _ptr(self.p_time64) if self._primary_src is None else None
        return A

    def _run(self, mode, **kw):
        A = self._args(mode, **kw)
        rc = _native.lib().rt_tmjp_run(ctypes.byref(A), _stream())
        _native.check(rc, 'rt_tmjp_run')

    # -- initial history (raoteh/sampler/_sample_tmjp_dense.py:509-627) ------------
    def initialize(self):
        """Primary trajectory first (0, 1, 3, 7, ... equally spaced events per edge until
        FFBS succeeds, raoteh/sampler/_sampler.py:612-643 via _sample_mcx_dense), then every
        tolerance class with one event at a uniform time in each primary segment."""
        DONE = 5
        k, j = 0, 0
        self.status.zero_()
        while True:
            if k > self.S:      # the reference's bound (raoteh/sampler/_sample_mcx_dense.py:100)
                raise RuntimeError('failed to find a feasible primary history')
            self._run(MODE_INIT_PRIMARY, init_k=k)
            failed = self.status == 1
            self.status[self.status == 0] = DONE
            if not bool(failed.any()):
                break
            self.status[failed] = 0
            k += 2 ** j
            j += 1
        self.status[self.status == DONE] = 0
        self.check()
        self._run(MODE_INIT_TOLERANCE)
        if bool((self.status == 6).any()):
            raise _native.NativeError('no feasible tolerance history: the disease data contradict '
                                      'the primary observations')
        self.check()
        self.sweeps_done = 1
        self.initialized = True
        return k

    def sweep(self, n_sweeps=1, stats=True, summary=False):
        """n_sweeps blocked Gibbs sweeps of every trajectory."""
        if not self.initialized:
            self.initialize()
        self.p_time64 = None          # the sampler state is float32 from here on
        flags = (F_STATS_PRIMARY | F_STATS_TOLERANCE if stats else 0)
        if summary:
            # the summary runs as its own launch after every sweep: with all warps of an SM in
            # the same pass the instruction working set stays small (measured: 9.8 ms per
            # sweep + summary of 6e4 C5 trajectories against 14.7 ms with RT_TMJP_F_SUMMARY)
            for _ in range(int(n_sweeps)):
                self._run(MODE_SWEEP, n_sweeps=1, flags=flags)
                self.sweeps_done += 1
                self._run(MODE_SUMMARY)
        else:
            self._run(MODE_SWEEP, n_sweeps=n_sweeps, flags=flags)
            self.sweeps_done += int(n_sweeps)
        self.check()

    def tolerance_summary(self):
        """get_tolerance_summary of every current primary trajectory -> [n_traj, 7]."""
        self._run(MODE_SUMMARY)
        if bool((self.status == 2).any()):
            from .sampler._util import NumericalZeroProb
            raise NumericalZeroProb('the denominator is zero')
        self.check()
        return self.summary_out[:, :7]

    def tolerance_log_likelihood(self, zero_as_neg_inf=False):
        """log-likelihood of every current primary trajectory under the compound process with
        the tolerance histories integrated out (raoteh/sampler/_tmjp.py:406-492,
        _tmjp_dense.py:407-505) -> [n_traj].  Runs the summary kernel.
        zero_as_neg_inf: a trajectory whose compound likelihood is numerically zero (status 2,
        where the reference raises NumericalZeroProb) gets -inf and a cleared status instead of
        an exception: as a Metropolis-Hastings target that is simply a rejected proposal."""
        if not zero_as_neg_inf:
            self.tolerance_summary()
            return self.summary_out[:, 7]
        self._run(MODE_SUMMARY)
        zero = self.status == 2
        self.status[zero] = 0
        self.check()
        return torch.where(zero, torch.full_like(self.summary_out[:, 7], float('-inf')),
                           self.summary_out[:, 7])

    def trajectory_log_likelihood(self):
        """log-likelihood of every current primary trajectory under the primary process itself
        (_mjp.get_trajectory_log_likelihood, raoteh/sampler/_mjp.py:186-250) -> [n_traj]."""
        self._run(MODE_TRAJ_LOGLIK)
        return self.traj_loglik

    def attach_primary(self, chains):
        """Use the primary trajectories of a raoteh_b200.raoteh.RaoTehChains (same tree, same
        trajectory count) for the summary / log-likelihood modes: Rao-Blackwellisation and
        importance weights for histories proposed under another rate matrix
        (raoteh/sampler/tests/test_sample_tmjp.py:186-246)."""
        if chains is not None and (chains.n_traj != self.n_traj or chains.sched.n != self.sched.n):
            raise ValueError('the attached chains must have the same tree and trajectory count')
        self._primary_src = chains

    def grow(self, cap_p=None, cap_t=None):
        """Re-allocate the jump / toggle pools with larger capacities (contents stay
        right-aligned) and clear the capacity flags.  The toggle pool of a class is limited to 255
        entries per trajectory (`t_total` and the per-branch counts are uint8 in the kernel's
        layout): a tree long enough to need more toggles of ONE class in one trajectory
        (tree length x omega_t of the order of 100) has to be cut into subtrees by the caller;
        status 3 after growing to 255 says so loudly rather than truncating."""
        T, dev, NP = self.n_traj, self.device, self.n_parts
        if cap_p is not None and int(cap_p) > self.cap_p:
            cap_p = int(cap_p)
            for name, dt in (('p_time', torch.float32), ('p_sb', torch.uint8)):
                new = torch.zeros((T, cap_p), dtype=dt, device=dev)
                new[:, cap_p - self.cap_p:] = getattr(self, name)
                setattr(self, name, new)
            self.cap_p = cap_p
        if cap_t is not None and int(cap_t) > self.cap_t:
            cap_t = min(255, int(cap_t))
            new = torch.zeros((T, NP, cap_t), dtype=torch.float32, device=dev)
            new[:, :, cap_t - self.cap_t:] = self.t_time
            self.t_time, self.cap_t = new, cap_t
        self.status[self.status == 3] = 0

    def check(self):
        st = self.status
        if int((st == 2).sum()):
            from .sampler._util import NumericalZeroProb
            raise NumericalZeroProb('the denominator is zero')
        if int((st == 4).sum()):
            raise _native.NativeError('a history has more real jumps than its capacity '
                                      '(cap_p=%d, cap_t=%d)' % (self.cap_p, self.cap_t))
        bad = int((st == 3).sum())
        if bad:
            raise _native.NativeError(
                '%d trajectories exceeded the event capacity (cap_p=%d, cap_t=%d) in one sweep; '
                're-create the sampler with larger capacities' % (bad, self.cap_p, self.cap_t))
        if int((st == 1).sum()) or int((st == 6).sum()):
            raise _native.NativeError('infeasible trajectory (structural zero): %d primary, %d tolerance'
                                      % (int((st == 1).sum()), int((st == 6).sum())))

    def reset_statistics(self):
        for t in (self.prim_dwell, self.prim_trans, self.tol_stats, self.summary_sum):
            t.zero_()

    # -- host-side views -------------------------------------------------------------
    def _edge_order(self):
        ops = self.ops.cpu().numpy()
        return [int(c) for code, c, a, b in ops if (code & 0xff) <= 2]

    def primary_trajectory(self, t):
        """(node states, dict child node -> (jump times from the parent end ascending,
        states of the k+1 segments parent side first))."""
        cnt = self.p_cnt[t].cpu().numpy().astype(int)
        tot = int(self.p_total[t])
        times = self.p_time[t, self.cap_p - tot:].cpu().numpy()
        sbs = self.p_sb[t, self.cap_p - tot:].cpu().numpy().astype(int)
        ns = self.p_node[t].cpu().numpy().astype(int)
        out, pos = {}, 0
        for c in self._edge_order():
            k = cnt[c]
            tt = times[pos:pos + k][::-1]
            sb = sbs[pos:pos + k][::-1]
            pos += k
            states = np.concatenate([sb, [ns[c]]]) if k else np.array([ns[c]])
            out[c] = (tt.astype(float), states)
        return ns, out

    def tolerance_trajectory(self, t, c):
        """Same view for tolerance class c (states 0 = off, 1 = on)."""
        cnt = self.t_cnt[t, :, c].cpu().numpy().astype(int)
        tot = int(self.t_total[t, c])
        times = self.t_time[t, c, self.cap_t - tot:].cpu().numpy()
        bits = (self.t_node[t].cpu().numpy().astype(np.int64) >> c) & 1
        out, pos = {}, 0
        for e in self._edge_order():
            k = cnt[e]
            tt = times[pos:pos + k][::-1]
            pos += k
            # the state alternates at every toggle; the child end carries bits[e]
            states = np.array([(bits[e] + (k - i)) % 2 for i in range(k + 1)], dtype=int)
            out[e] = (tt.astype(float), states)
        return bits, out

    def load_primary_trajectories(self, node_states, edge_jumps):
        """Install caller-given primary trajectories (for the summary of histories sampled
        elsewhere).  node_states: int [n_traj, n_nodes]; edge_jumps: list over trajectories of
        dict child node -> (times from the parent end ascending, parent-side states)."""
        T, n = self.n_traj, self.sched.n
        order = self._edge_order()
        p_cnt = np.zeros((T, n), dtype=np.uint8)
        p_time = np.zeros((T, self.cap_p), dtype=np.float64)
        p_sb = np.zeros((T, self.cap_p), dtype=np.uint8)
        p_total = np.zeros(T, dtype=np.int32)
        for t in range(T):
            tt_all, sb_all = [], []
            for c in order:
                times, sbs = edge_jumps[t].get(c, ((), ()))
                p_cnt[t, c] = len(times)
                tt_all.extend(list(times)[::-1])      # child end first
                sb_all.extend(list(sbs)[::-1])
            k = len(tt_all)
            if k > self.cap_p:
                raise ValueError('trajectory has more jumps than cap_p')
            p_total[t] = k
            if k:
                p_time[t, self.cap_p - k:] = tt_all
                p_sb[t, self.cap_p - k:] = sb_all
        dev = self.device
        self.p_node.copy_(torch.from_numpy(np.asarray(node_states, dtype=np.uint8)).to(dev))
        self.p_cnt.copy_(torch.from_numpy(p_cnt).to(dev))
        self.p_time.copy_(torch.from_numpy(p_time.astype(np.float32)).to(dev))
        self.p_time64 = torch.from_numpy(p_time).to(dev)      # the summary uses the exact times
        self.p_sb.copy_(torch.from_numpy(p_sb).to(dev))
        self.p_total.copy_(torch.from_numpy(p_total).to(dev))
        self.status.zero_()
