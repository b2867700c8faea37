"""
Batched Metropolis-Hastings on top of Rao-Teh proposals, entirely on the device.

The reference's `_sampler.gen_mh_histories` (raoteh/sampler/_sampler.py:393-551) proposes a
history with one Rao-Teh sweep under a rate matrix Q and accepts it with probability
min(1, target(new) proposal(old) / (target(old) proposal(new))), where `proposal` is the
trajectory likelihood under Q and `target` is a caller-supplied density.  Its use in the
tolerance experiments (raoteh/sampler/tests/test_sample_tmjp.py:248-276) has
target = `_tmjp.get_tolerance_process_log_likelihood`: primary histories are proposed under the
approximate primary process (`get_primary_proposal_rate_matrix`) and corrected to the primary
marginal of the true compound process.  `ToleranceMetropolisChains` runs exactly that for every
(chain, site) at once: proposals by rt_raoteh_sweeps, both log-likelihoods by rt_tmjp_run
(modes RT_TMJP_TRAJ_LOGLIK and RT_TMJP_SUMMARY), accept/reject and the roll-back of rejected
trajectories as masked copies of the trajectory tensors.
"""
from __future__ import annotations

import numpy as np
import torch

from .raoteh import RaoTehChains
from .tmjp import ToleranceChains


class ToleranceMetropolisChains(object):
    def __init__(self, sched, Q_primary, primary_distn, primary_to_part, rate_on, rate_off, obs,
                 n_chains=1, uniformization_factor=2.0, cap=96, seed=0, device='cuda',
                 tol_obs=None, tol_obs_nodes=None):
        from .sampler._tmjp_dense import get_primary_proposal_rate_matrix, get_two_state_tolerance_distn
        self.S = S = int(np.asarray(Q_primary).shape[0])
        part = dict((s, int(primary_to_part[s])) for s in range(S))
        tol_distn = get_two_state_tolerance_distn(rate_off, rate_on)
        self.Q_proposal = get_primary_proposal_rate_matrix(np.asarray(Q_primary, dtype=float), part, tol_distn)
        self.proposal = RaoTehChains(sched, self.Q_proposal, obs, n_chains=n_chains,
                                     root_distn=primary_distn, cap=cap, seed=seed,
                                     uniformization_factor=uniformization_factor, device=device)
        self.target = ToleranceChains(sched, Q_primary, primary_distn, part, rate_on, rate_off, obs,
                                      n_chains=n_chains, cap_p=cap, cap_t=8, seed=seed, device=device,
                                      tol_obs=tol_obs, tol_obs_nodes=tol_obs_nodes)
        self.target.attach_primary(self.proposal)
        self.gen = torch.Generator(device=self.proposal.device)
        self.gen.manual_seed(int(seed) + 12345)
        self.n_traj = self.proposal.n_traj
        self.ll_biased = self.ll_target = None
        self.n_accepted = 0
        self.n_proposed = 0
        # Rao-Blackwellised tolerance summary (get_tolerance_summary, 7 values) of the CURRENT
        # history of every chain, and its sum over chains and steps
        self.cur_summary = None
        self.summary_acc = torch.zeros(7, dtype=torch.float64, device=self.proposal.device)

    @property
    def dwell_sum(self):
        return self.proposal.dwell_sum

    @property
    def trans_sum(self):
        return self.proposal.trans_sum

    def initialize(self):
        self.proposal.initialize()
        self.ll_biased = self.proposal.trajectory_log_likelihood().clone()
        self.ll_target = self.target.tolerance_log_likelihood().clone()
        self.cur_summary = self.target.summary_out[:, :7].clone()

    def step(self, n_steps=1, stats=True):
        """n_steps proposals per trajectory; with stats=True the statistics of the CURRENT
        history (accepted proposal or repeated previous one) are added after every step."""
        if self.ll_biased is None:
            self.initialize()
        for _ in range(int(n_steps)):
            saved = self.proposal.snapshot()
            self.proposal.sweep(1, stats=False)
            ll_b = self.proposal.trajectory_log_likelihood()
            ll_t = self.target.tolerance_log_likelihood(zero_as_neg_inf=True)   # zero target = rejection
            log_ratio = ll_t - self.ll_target - ll_b + self.ll_biased          # _sampler.py:514-518
            u = torch.rand(self.n_traj, dtype=torch.float64, device=log_ratio.device, generator=self.gen)
            accept = torch.log(u) < log_ratio
            self.proposal.restore(saved, reject=~accept)
            self.ll_biased = torch.where(accept, ll_b, self.ll_biased)
            self.ll_target = torch.where(accept, ll_t, self.ll_target)
            self.cur_summary = torch.where(accept[:, None], self.target.summary_out[:, :7], self.cur_summary)
            self.n_accepted += int(accept.sum())
            self.n_proposed += self.n_traj
            if stats:
                self.proposal.trajectory_log_likelihood(stats=True)
                self.summary_acc += self.cur_summary.sum(dim=0)
