// K6: Rao-Teh uniformization sweeps on trees, one thread per (chain, site)
// trajectory; counter-based Philox4x32-10 per (trajectory, sweep), which seeds a PCG32 stream for
// the events inside the sweep.
//
// Replaces the reference's per-trajectory Python/networkx sweep
//   _sampler.gen_restricted_histories loop      raoteh/sampler/_sampler.py:366-390
//   _sample_mjp.resample_poisson                 raoteh/sampler/_sample_mjp.py:19-69
//   _graph_transform.get_chunk_tree_type_b       raoteh/sampler/_graph_transform.py:298
//   _sample_mcy.resample_edge_states -> FFBS     raoteh/sampler/_sample_mcy.py:86-187,
//                                                _sample_mc0.py:20-93
//   _graph_transform.remove_redundant_nodes      raoteh/sampler/_graph_transform.py:144
//   _mjp.get_history_statistics                  raoteh/sampler/_mjp.py:150 (dwell :74, transitions :97)
//   _sampler.get_restricted_feasible_history     raoteh/sampler/_sampler.py:563-643 (init mode)
//
// A trajectory is stored implicitly: a state per tree node plus, per edge, the
// list of real jumps (time from the parent end, state on the parent side).  The
// chunk tree of the reference is never built: an edge with k events is a chain
// of k steps of the uniformized matrix B = I + Q/omega between its end points,
// and an edge without events forces equal end states.
//
// One sweep = two passes over the host-built upward program (the same program
// the pruning kernels walk):
//   UP   (child -> parent along every edge): thin a Poisson process of rate
//        omega - q_s onto every constant-state segment, and at each event
//        (virtual or old jump) record the backward message just below it and push
//        it through B.  Only events need a stored message: an edge without events
//        forces equal end states, so node partials live only in a small per-thread
//        slot stack in shared memory.  Virtual events are drawn in hazard space
//        (unit-rate exponential gaps consumed by rate*length of each segment), so
//        a segment without an event costs one multiply-compare and no random number.
//   DOWN (program reversed): sample the root from pi * L_root, then at every event
//        draw the child-side state from B[parent state, :] * message, drop
//        self-transitions, accumulate dwell times and transition counts, and
//        rewrite the jump lists.  This is FFBS on the implicit chunk tree.
#include "rt_common.cuh"
#include "rt_philox.cuh"
#include "../../include/rt_b200.h"

namespace {

constexpr int kBlock = 128;

template <int S, int OBS>
__device__ __forceinline__ void load_obs_vec(const void* obs, int slot, int64_t obs_stride,
                                             int64_t site, double (&v)[S]) {
  if (OBS == OBS_CODES) {
    const int k = reinterpret_cast<const uint8_t*>(obs)[(int64_t)slot * obs_stride + site];
#pragma unroll
    for (int s = 0; s < S; ++s) v[s] = (k == RT_MISSING || k == s) ? 1.0 : 0.0;
  } else {
    const unsigned long long mk =
        reinterpret_cast<const unsigned long long*>(obs)[(int64_t)slot * obs_stride + site];
#pragma unroll
    for (int s = 0; s < S; ++s) v[s] = ((mk >> s) & 1ull) ? 1.0 : 0.0;
  }
}

// TT = type of every event time, branch position and hazard: float (default; 4-byte jump
// lists) or double (the reference's arithmetic; rt_raoteh_sweeps_f64)
template <typename TT> struct TimeOps;
template <> struct TimeOps<float> {
  static __device__ __forceinline__ float neglog_unit(uint32_t u) { return -__logf(RT_U32_TO_UNIT(u)); }
};
template <> struct TimeOps<double> {
  static __device__ __forceinline__ double neglog_unit(uint32_t u) {
    return -log(((double)u + 0.5) * 2.3283064365386963e-10);
  }
};

template <typename TT>
struct SweepArgs {
  int n_nodes, n_ops, n_slots, cap, scr_cap;
  int64_t n_traj, stride, n_sites, obs_stride, traj0;
  int64_t scr_stride;    // trajectories per launch chunk (row length of scr_count)
  const int4* program;
  const int32_t* parent;
  const double* length;
  const double* B;
  const double* rate;
  const double* root_distn;
  const void* obs;
  uint8_t* node_state;   // [n_nodes][stride]
  TT* ev_time;           // [stride][cap]   jumps in up order, occupying [cap - total, cap)
  uint8_t* ev_sb;        // [stride][cap]   state on the parent side of the jump
  uint8_t* ev_count;     // [n_nodes][stride]
  int32_t* ev_total;     // [stride]
  // candidate events of the sweep, up order, one contiguous record per event and
  // trajectory of the launch chunk: [scr_stride][scr_cap] x { double beta[S]; TT time; uint32 u; }
  // padded to 16 B
  unsigned char* scr_rec;
  uint8_t* scr_count;    // [n_nodes][scr_stride]
  unsigned long long seed;
  long long sweep0;
  int n_sweeps;
  int init_k;            // >= 0: initial-history mode with init_k equally spaced events per edge
  int32_t* sweep_count;  // nullable [stride]: sweeps completed per trajectory (see rt_raoteh_args)
  double* dwell_sum;     // [S]   += over trajectories and sweeps
  double* trans_sum;     // [S*S] +=
  int8_t* status;        // [stride]
};

// S <= 4: 80 registers, 6 CTAs per SM (the shared-memory limit): the kernel is latency bound and
// 24 warps per SM beat 16 by 9 % at C4 (profiles/r1_raoteh_occupancy.log; 5 CTAs: no gain)
template <int S, int OBS, bool STATS, typename TT>
__global__ void __launch_bounds__(kBlock, (S <= 4 && sizeof(TT) == 4 ? 6 : 4))
raoteh_kernel(SweepArgs<TT> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int4* prog_s = reinterpret_cast<int4*>(smem_raw);
  double* B_s = reinterpret_cast<double*>(prog_s + A.n_ops);
  double* pi_s = B_s + S * S;
  double* dwell_s = pi_s + S;                                       // [S][kBlock]
  uint32_t* trans_s = reinterpret_cast<uint32_t*>(dwell_s + S * kBlock);   // [S*S][kBlock]
  TT* rate_s = reinterpret_cast<TT*>(trans_s + S * S * kBlock);      // [S] poisson rates
  TT* rinv_s = rate_s + S;                                       // [S]
  TT* len_s = rinv_s + S;                                        // [n_nodes]
  int* par_s = reinterpret_cast<int*>(len_s + A.n_nodes);           // [n_nodes]
  double* stk = reinterpret_cast<double*>(
      (reinterpret_cast<uintptr_t>(par_s + A.n_nodes) + 7) & ~(uintptr_t)7);   // [n_slots][S][kBlock]
  __shared__ double red[kBlock / 32][S * S + S];

  const int tid = threadIdx.x;
  for (int i = tid; i < A.n_ops; i += kBlock) prog_s[i] = A.program[i];
  for (int i = tid; i < S * S; i += kBlock) B_s[i] = A.B[i];
  if (tid < S) {
    rate_s[tid] = (TT)A.rate[tid];
    rinv_s[tid] = A.rate[tid] > 0.0 ? (TT)(1.0 / A.rate[tid]) : (TT)0;
    pi_s[tid] = A.root_distn ? A.root_distn[tid] : 1.0;
  }
  for (int i = tid; i < A.n_nodes; i += kBlock) { len_s[i] = (TT)A.length[i]; par_s[i] = A.parent[i]; }
#pragma unroll
  for (int s = 0; s < S; ++s) dwell_s[s * kBlock + tid] = 0.0;
#pragma unroll
  for (int q = 0; q < S * S; ++q) trans_s[q * kBlock + tid] = 0u;
  __syncthreads();

  const int64_t traj = (int64_t)blockIdx.x * kBlock + tid;
  const bool active = traj < A.n_traj && A.status[traj] == RT_SITE_OK;
  const int64_t site = active ? (A.traj0 + traj) % A.n_sites : 0;
  const int64_t st = A.stride;

  if (active) {
    Philox rng;
    uint8_t* ns_p = A.node_state + traj;
    uint8_t* cnt_p = A.ev_count + traj;
    uint8_t* scnt_p = A.scr_count + traj;
    const int64_t sst = A.scr_stride;
    TT* evt_p = A.ev_time + (size_t)traj * A.cap;      // [traj][cap]: a thread's jumps are contiguous
    uint8_t* evs_p = A.ev_sb + (size_t)traj * A.cap;
    constexpr int kRec = (8 * S + (int)sizeof(TT) + 4 + 15) / 16 * 16;        // bytes per scratch record
    unsigned char* rec_p = A.scr_rec + (size_t)traj * A.scr_cap * kRec;
    // with a per-trajectory counter the trajectory resumes at its own sweep index
    const long long sw_end = A.sweep0 + A.n_sweeps;
    long long sw = (A.sweep_count && A.init_k < 0) ? (long long)A.sweep_count[traj] : A.sweep0;
    for (; sw < sw_end; ++sw) {
      rng.init(A.seed, (uint64_t)(A.traj0 + traj), (uint32_t)sw);
      // substream 0: root draw (word 0) and the first unit-rate gap (word 1)
      rng.block(0u, 0u);
      const uint32_t u_root = rng.out[0];
      // Words 2 and 3 of the sweep's Philox block seed a PCG32 (XSH-RR, O'Neill 2014) stream for
      // the events of this sweep: the Philox block keeps the draw addressable by (seed, trajectory,
      // sweep) -- results do not depend on launch partitioning or rank count -- while the per-event
      // cost drops from a ten-round Philox block per edge (~65 instructions, executed by the whole
      // warp whenever ANY lane starts a block) to ~12 instructions per word.
      uint64_t pcg_state = ((uint64_t)rng.out[3] << 32) | (uint64_t)rng.out[2];
      const uint64_t pcg_inc = ((uint64_t)(A.traj0 + traj) << 1) | 1ull;
      auto pcg_next = [&]() -> uint32_t {
        const uint64_t old = pcg_state;
        pcg_state = old * 6364136223846793005ull + pcg_inc;
        const uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
        const uint32_t rot = (uint32_t)(old >> 59);
        return (xs >> rot) | (xs << ((32u - rot) & 31u));
      };
      pcg_next();      // one step away from the raw seed
      // Virtual events: a Poisson process of rate omega - q_s on every segment, sampled
      // in HAZARD space.  Walking the segments in program order, `hrem` is the hazard left
      // until the next event (unit-rate exponential gaps); a segment of hazard h consumes
      // it, and only an actual event costs a random number and a logarithm.
      TT hrem = TimeOps<TT>::neglog_unit(rng.out[1]);
      // =============================== UP ===============================
      int nA = 0;                                   // entries pushed to the scratch list
      int rd = A.cap - A.ev_total[traj];            // read cursor in the old jump list
      // the old jump list is consumed strictly in order (across edges too), so its next entry
      // is loaded when the previous one is consumed, long before it is needed
      TT pf_time = (TT)-1;
      int pf_sb = 0;
      if (A.init_k < 0 && rd < A.cap) { pf_time = evt_p[rd]; pf_sb = evs_p[rd]; }
      bool overflow = false;
      double acc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s] = 1.0;
      int root_state = 0;
      bool infeasible = false;

      for (int ip = 0; ip < A.n_ops; ++ip) {
        const int4 op = prog_s[ip];
        const int code = op.x & 0xff;
        if (code <= OP_MSG_ONES) {
          const int c = op.y;
          double beta[S];
          if (code == OP_MSG_SLOT) {
#pragma unroll
            for (int s = 0; s < S; ++s) beta[s] = stk[(op.z * S + s) * kBlock + tid];
          } else if (code == OP_MSG_OBS) {
            load_obs_vec<S, OBS>(A.obs, op.z, A.obs_stride, site, beta);
          } else {
#pragma unroll
            for (int s = 0; s < S; ++s) beta[s] = 1.0;
          }
          const TT tc = len_s[c];
          // --- walk the edge from the child end to the parent end over its candidate
          //     events (old jumps + fresh virtual events): one loop, one event site ---
          int kA = 0;
          int cur = 0, k_old = 0;
          TT next_old = (TT)-1;            // time of the next old jump toward the parent (-1: none)
          int next_sb = 0;
          TT pos = tc;                     // current position (moving toward 0)
          int init_left = 0;
          if (A.init_k >= 0) {
            // no events on a zero-length branch: its end states are equal (P(0) = I), and a jump
            // at time 0 would be indistinguishable from 'no more jumps' in the sweeps
            init_left = tc > (TT)0 ? A.init_k : 0;
          } else {
            cur = ns_p[(int64_t)c * st];
            k_old = cnt_p[(int64_t)c * st];
            if (k_old > 0) {
              next_old = pf_time;
              next_sb = pf_sb;
              if (rd + 1 < A.cap) { pf_time = evt_p[rd + 1]; pf_sb = evs_p[rd + 1]; }
            }
          }
          while (true) {
            TT cand;                       // position of the next candidate event
            bool is_old = false, is_virtual = false;
            if (A.init_k >= 0) {
              if (init_left == 0) break;
              cand = tc * (TT)init_left / (TT)(A.init_k + 1);
              --init_left;
            } else {
              const TT seg_start = next_old > (TT)0 ? next_old : (TT)0;
              const TT h = rate_s[cur] * (pos - seg_start);
              if (hrem < h) {
                cand = pos - hrem * rinv_s[cur];
                // strictly inside the segment: a virtual event that rounds onto an end of its
                // segment would duplicate an existing jump time (or the branch end)
                if (!(cand > seg_start && cand < pos)) cand = seg_start + (TT)0.5 * (pos - seg_start);
                is_virtual = true;
              } else {
                hrem -= h;
                if (next_old > (TT)0) { cand = next_old; is_old = true; }
                else break;
              }
            }
            // random words of this event: the sweep's sequential stream (see pcg_next)
            const uint32_t u_draw = pcg_next();
            if (is_virtual) hrem = TimeOps<TT>::neglog_unit(pcg_next());
            // record beta just below the event, then push it through B
            if (nA < A.scr_cap) {
              unsigned char* rp = rec_p + (size_t)nA * kRec;
              if constexpr (S % 2 == 0) {
#pragma unroll
                for (int s = 0; s < S; s += 2)
                  reinterpret_cast<double2*>(rp)[s / 2] = make_double2(beta[s], beta[s + 1]);
              } else {
#pragma unroll
                for (int s = 0; s < S; ++s) reinterpret_cast<double*>(rp)[s] = beta[s];
              }
              if constexpr (sizeof(TT) == 4) {
                *reinterpret_cast<float2*>(rp + 8 * S) = make_float2((float)cand, __uint_as_float(u_draw));
              } else {
                *reinterpret_cast<double*>(rp + 8 * S) = (double)cand;
                *reinterpret_cast<uint32_t*>(rp + 8 * S + 8) = u_draw;
              }
            } else overflow = true;
            ++nA; ++kA;
            {
              double nb[S];
#pragma unroll
              for (int a2 = 0; a2 < S; ++a2) {
                double t2 = 0.0;
#pragma unroll
                for (int s = 0; s < S; ++s) t2 = fma(B_s[a2 * S + s], beta[s], t2);
                nb[a2] = t2;
              }
#pragma unroll
              for (int s = 0; s < S; ++s) beta[s] = nb[s];
            }
            pos = cand;
            if (is_old) {
              cur = next_sb;
              ++rd;
              --k_old;
              if (k_old > 0) {
                next_old = pf_time;
                next_sb = pf_sb;
                if (rd + 1 < A.cap) { pf_time = evt_p[rd + 1]; pf_sb = evs_p[rd + 1]; }
              } else {
                next_old = (TT)-1;
              }
            }
          }
          if (kA > 255) overflow = true;
          scnt_p[(int64_t)c * sst] = (uint8_t)(kA > 255 ? 255 : kA);
          if (kA > 0) {   // keep the chain of B-steps in range
            double mx = beta[0];
#pragma unroll
            for (int s = 1; s < S; ++s) mx = fmax(mx, beta[s]);
            if (mx > 0.0) {
              const double sc = rt_pow2_neg(rt_exponent(mx));
#pragma unroll
              for (int s = 0; s < S; ++s) beta[s] *= sc;
            }
          }
#pragma unroll
          for (int s = 0; s < S; ++s) acc[s] *= beta[s];
        } else if (code == OP_APPLY_OBS) {
          double v[S];
          load_obs_vec<S, OBS>(A.obs, op.z, A.obs_stride, site, v);
#pragma unroll
          for (int s = 0; s < S; ++s) acc[s] *= v[s];
        } else {   // OP_STORE / OP_ROOT
          double mx = acc[0];
#pragma unroll
          for (int s = 1; s < S; ++s) mx = fmax(mx, acc[s]);
          if (mx > 0.0) {
            const double sc = rt_pow2_neg(rt_exponent(mx));
#pragma unroll
            for (int s = 0; s < S; ++s) acc[s] *= sc;
          }
          if (code == OP_STORE) {
#pragma unroll
            for (int s = 0; s < S; ++s) { stk[(op.z * S + s) * kBlock + tid] = acc[s]; acc[s] = 1.0; }
          } else {
            // root ~ pi * L_root  (_sample_mc0.py:41-90)
            double w[S], tot = 0.0;
#pragma unroll
            for (int s = 0; s < S; ++s) { w[s] = pi_s[s] * acc[s]; tot += w[s]; }
            if (!(tot > 0.0)) infeasible = true;
            const double x = ((double)u_root + 0.5) * 2.3283064365386963e-10 * tot;
            double cum = 0.0;
            bool found = false;
#pragma unroll
            for (int s = 0; s < S; ++s) {
              cum += w[s];
              if (!found && w[s] > 0.0 && x <= cum) { root_state = s; found = true; }
            }
            if (!found) {
#pragma unroll
              for (int s = 0; s < S; ++s) if (w[s] > 0.0) root_state = s;
            }
          }
        }
      }
      if (infeasible) { A.status[traj] = RT_SITE_STRUCTURAL_ZERO; break; }
      if (overflow) { A.status[traj] = 3; break; }

      // ============================== DOWN ==============================
      ns_p[0] = (uint8_t)root_state;
      int rdA = nA;          // scratch is consumed backwards
      int wr = A.cap;        // new jump list grows backwards from the end
      // ... strictly in order, so record rdA - 2 is loaded while record rdA - 1 is sampled
      TT t_n = (TT)0;
      uint32_t u_n = 0u;
      double b_n[S];
#pragma unroll
      for (int s = 0; s < S; ++s) b_n[s] = 0.0;
      auto load_rec = [&](int idx) {
        const unsigned char* rp = rec_p + (size_t)idx * kRec;
        if constexpr (S % 2 == 0) {
#pragma unroll
          for (int s = 0; s < S; s += 2) {
            const double2 v = reinterpret_cast<const double2*>(rp)[s / 2];
            b_n[s] = v.x; b_n[s + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int s = 0; s < S; ++s) b_n[s] = reinterpret_cast<const double*>(rp)[s];
        }
        if constexpr (sizeof(TT) == 4) {
          const float2 tu = *reinterpret_cast<const float2*>(rp + 8 * S);
          t_n = (TT)tu.x; u_n = __float_as_uint(tu.y);
        } else {
          t_n = (TT)*reinterpret_cast<const double*>(rp + 8 * S);
          u_n = *reinterpret_cast<const uint32_t*>(rp + 8 * S + 8);
        }
      };
      if (nA > 0) load_rec(nA - 1);
      bool pool_overflow = false;
      for (int ip = A.n_ops - 1; ip >= 0; --ip) {
        const int4 op = prog_s[ip];
        if ((op.x & 0xff) > OP_MSG_ONES) continue;
        const int c = op.y;
        int cur = ns_p[(int64_t)par_s[c] * st];
        const int kA = scnt_p[(int64_t)c * sst];
        const TT tc = len_s[c];
        TT prev = (TT)0;
        int kept = 0;
        for (int j = 0; j < kA; ++j) {
          --rdA;
          const TT tau = t_n;
          const uint32_t u_draw = u_n;
          // child-side state ~ B[cur, :] * beta  (_sample_mc0.py:66-90)
          double w[S], tot = 0.0;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            w[s] = B_s[cur * S + s] * b_n[s];
            tot += w[s];
          }
          if (rdA > 0) load_rec(rdA - 1);
          const double x = ((double)u_draw + 0.5) * 2.3283064365386963e-10 * tot;
          double cum = 0.0;
          int nxt = -1;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            cum += w[s];
            if (nxt < 0 && w[s] > 0.0 && x <= cum) nxt = s;
          }
          if (nxt < 0) {
#pragma unroll
            for (int s = 0; s < S; ++s) if (w[s] > 0.0) nxt = s;
          }
          if (STATS) dwell_s[cur * kBlock + tid] += (double)(tau - prev);
          prev = tau;
          if (nxt != cur) {
            if (STATS) trans_s[(cur * S + nxt) * kBlock + tid] += 1u;
            --wr;
            if (wr >= 0) {
              evt_p[wr] = tau;
              evs_p[wr] = (uint8_t)cur;
            } else pool_overflow = true;
            ++kept;
            cur = nxt;
          }
        }
        if (STATS) dwell_s[cur * kBlock + tid] += (double)(tc - prev);
        ns_p[(int64_t)c * st] = (uint8_t)cur;
        cnt_p[(int64_t)c * st] = (uint8_t)kept;
      }
      // jumps were written in down order from the end backwards == up order forwards
      if (pool_overflow) { A.status[traj] = 4; A.ev_total[traj] = 0; break; }
      A.ev_total[traj] = A.cap - wr;
    }
    // sw = first sweep NOT completed (== sw_end unless the trajectory stopped early)
    if (A.sweep_count) A.sweep_count[traj] = (int32_t)sw;
  }

  if (STATS && A.dwell_sum) {
#pragma unroll
    for (int q = 0; q < S + S * S; ++q) {
      const double mine = q < S ? dwell_s[(q < S ? q : 0) * kBlock + tid]
                                : (double)trans_s[(q >= S ? q - S : 0) * kBlock + tid];
      const double v = rt_warp_sum(mine);
      if ((tid & 31) == 0) red[tid >> 5][q] = v;
    }
    __syncthreads();
    if (tid < S + S * S) {
      double t = 0.0;
      for (int i = 0; i < kBlock / 32; ++i) t += red[i][tid];
      if (t != 0.0) atomicAdd(tid < S ? &A.dwell_sum[tid] : &A.trans_sum[tid - S], t);
    }
  }
}

template <int S, int OBS, typename TT>
int launch(const SweepArgs<TT>& A, bool stats, cudaStream_t stream) {
  size_t smem = sizeof(int4) * A.n_ops + sizeof(double) * (S * S + S) +
                sizeof(double) * S * kBlock + sizeof(uint32_t) * S * S * kBlock +
                sizeof(TT) * 2 * S + sizeof(TT) * A.n_nodes + sizeof(int) * A.n_nodes +
                sizeof(double) * (size_t)A.n_slots * S * kBlock + 32;
  if (smem > 200 * 1024) return RT_ERR_UNSUPPORTED;
  const unsigned grid = (unsigned)((A.n_traj + kBlock - 1) / kBlock);
  if (stats) {
    auto kern = raoteh_kernel<S, OBS, true, TT>;
    RT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kBlock, smem, stream>>>(A);
  } else {
    auto kern = raoteh_kernel<S, OBS, false, TT>;
    RT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kBlock, smem, stream>>>(A);
  }
  RT_CUDA_CHECK(cudaGetLastError());
  return RT_OK;
}

template <int S, typename TT>
int launch_s(int obs_kind, const SweepArgs<TT>& A, bool stats, cudaStream_t stream) {
  if (obs_kind == OBS_CODES) return launch<S, OBS_CODES, TT>(A, stats, stream);
  if (obs_kind == OBS_MASK) return launch<S, OBS_MASK, TT>(A, stats, stream);
  return RT_ERR_ARG;
}

// Trajectories per launch: the per-event scratch (cap records of 48 B at S = 4) is sized for one
// chunk and reused by the next launch on the same stream, so the workspace does not grow with the
// number of trajectories (4096 chains x 1e4 sites = 4.1e7 trajectories at BASELINE config C4).
// 2^19 trajectories = 4.6 waves of 6 CTAs x 148 SMs.
constexpr int64_t kChunkTraj = 1 << 19;

template <typename TT>
int dispatch_t(const rt_raoteh_args& R, cudaStream_t stream) {
  const int S = R.S, obs_kind = R.obs_kind, n_nodes = R.n_nodes, n_ops = R.n_ops, n_slots = R.n_slots;
  const int cap = R.cap, n_sweeps = R.n_sweeps, init_k = R.init_k;
  const int64_t n_traj = R.n_traj, stride = R.traj_stride, n_sites = R.n_sites, traj0 = R.traj0;
  const int64_t obs_stride = R.obs_stride, sweep0 = R.sweep0;
  const uint64_t seed = R.seed;
  const int32_t* program = R.program; const int32_t* parent = R.parent;
  const double* length = R.length; const double* B = R.B; const double* rate = R.rate;
  const double* root_distn = R.root_distn; const void* obs = R.obs;
  uint8_t* node_state = R.node_state; TT* ev_time = reinterpret_cast<TT*>(R.ev_time);
  uint8_t* ev_sb = R.ev_sb; uint8_t* ev_count = R.ev_count; int32_t* ev_total = R.ev_total;
  double* dwell_sum = R.dwell_sum; double* trans_sum = R.trans_sum; int8_t* status = R.status;
  SweepArgs<TT> A;
  A.n_nodes = n_nodes; A.n_ops = n_ops; A.n_slots = n_slots; A.cap = cap;
  // sweeps: kept jumps <= scratch entries <= cap, so the jump list cannot overflow.
  // init mode: init_k events on every edge are candidates (status 4 if more than
  // `cap` of them turn out to be real jumps).
  A.scr_cap = cap;
  if (init_k > 0 && (n_nodes - 1) * init_k > cap) A.scr_cap = (n_nodes - 1) * init_k;
  A.stride = stride; A.n_sites = n_sites; A.obs_stride = obs_stride;
  A.program = reinterpret_cast<const int4*>(program);
  A.parent = parent; A.length = length; A.B = B; A.rate = rate; A.root_distn = root_distn;
  A.obs = obs;
  A.seed = seed; A.sweep0 = sweep0; A.n_sweeps = n_sweeps; A.init_k = init_k;
  A.dwell_sum = dwell_sum; A.trans_sum = trans_sum;
  const int64_t chunk = n_traj < kChunkTraj ? n_traj : kChunkTraj;
  A.scr_stride = chunk;
  // scratch of one chunk: event records and per-edge candidate counts
  unsigned char* ws = nullptr;
  const size_t n_scr = (size_t)A.scr_cap * (size_t)chunk;
  const size_t rec = (size_t)((8 * S + (int)sizeof(TT) + 4 + 15) / 16 * 16);
  const size_t bytes = n_scr * rec + (size_t)n_nodes * (size_t)chunk + 64;
  RT_CUDA_CHECK(rt_ws_alloc((void**)&ws, bytes, stream));
  A.scr_rec = ws;
  A.scr_count = ws + n_scr * rec;
  const bool stats = dwell_sum != nullptr && trans_sum != nullptr && init_k < 0;
  int rc = RT_OK;
  for (int64_t lo = 0; lo < n_traj && rc == RT_OK; lo += chunk) {
    A.n_traj = (n_traj - lo < chunk) ? n_traj - lo : chunk;
    A.traj0 = traj0 + lo;
    A.node_state = node_state + lo; A.ev_count = ev_count + lo; A.ev_total = ev_total + lo;
    A.ev_time = ev_time + (size_t)lo * cap; A.ev_sb = ev_sb + (size_t)lo * cap;
    A.status = status + lo;
    A.sweep_count = R.sweep_count ? R.sweep_count + lo : nullptr;
    switch (S) {
      case 2: rc = launch_s<2, TT>(obs_kind, A, stats, stream); break;
      case 3: rc = launch_s<3, TT>(obs_kind, A, stats, stream); break;
      case 4: rc = launch_s<4, TT>(obs_kind, A, stats, stream); break;
      case 5: rc = launch_s<5, TT>(obs_kind, A, stats, stream); break;
      case 6: rc = launch_s<6, TT>(obs_kind, A, stats, stream); break;
      case 8: rc = launch_s<8, TT>(obs_kind, A, stats, stream); break;
      default: rc = RT_ERR_UNSUPPORTED;
    }
  }
  rt_ws_free(ws, stream);
  return rc;
}

}  // namespace

int rt_raoteh_dispatch(const rt_raoteh_args& R, cudaStream_t stream) {
  return R.time_f64 ? dispatch_t<double>(R, stream) : dispatch_t<float>(R, stream);
}
