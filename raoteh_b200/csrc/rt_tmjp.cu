// K6w / K7 / K8: warp-cooperative Rao-Teh kernels, one warp per (chain, site) trajectory.
//
//   primary_pass  Rao-Teh sweep of an S <= 64 state trajectory; lanes own the states of
//                 the FFBS messages (matvec and inverse-CDF draw are warp collectives),
//                 the walk over edges and events is warp-uniform (no divergence).
//                 With tolerance trajectories present, a chunk's allowed primary states
//                 are those whose class is never OFF inside the chunk:
//                   raoteh/sampler/_sample_tmjp_dense.py:175-371 (resample_primary_states_v1)
//                 Without (n_parts == 0) it is the plain sweep of
//                   raoteh/sampler/_sampler.py:366-390 for 9 <= S <= 64.
//   tol_pass      lane = tolerance class: Poisson events + 2-state FFBS given the primary
//                 trajectory, class forced ON wherever the primary state belongs to it,
//                 disease data at nodes:
//                   raoteh/sampler/_sample_tmjp_dense.py:146-169, :374-506
//                 init mode: one event at a uniform time in every primary segment (:567-611).
//   summary_pass  lane = tolerance class: Rao-Blackwellised tolerance expectations of a
//                 primary trajectory, closed-form 2x2 block of expm(t Q3) and its Frechet
//                 integrals per segment:
//                   raoteh/sampler/_tmjp_dense.py:724-855 (get_tolerance_summary),
//                   :965-1078 (get_inhomogeneous_mjp), :246-349, _linalg.py:14-118.
//
// The merged trees / chunk trees of the reference (_graph_transform.py:298,508) are never
// built: per edge, the walk merges the jump lists on the fly.
#include "rt_common.cuh"
#include "rt_philox.cuh"
#include "../../include/rt_b200.h"

namespace {

// warps (= trajectories in flight) per CTA: 16 when the per-warp shared-memory state fits
// (measured best on C5: 16 warps per SM at 128 registers), else 8, 4 or 2; the register
// budget is always sized for 16 warps per SM.
constexpr int kMaxWarps = 16;
constexpr unsigned FULL = 0xffffffffu;

struct Scratch {
  double* p_beta;      // [n_warps][scr_cap][SPAD]    primary: message just below every event
  float2* t_tu;        // [n_warps][cap_ts][32]       tolerance: (time, uniform) per event and lane
  double2* t_beta;     // [n_warps][cap_ts][32]
  double2* seg;        // [n_warps][n_seg][32]        summary: message at the low end of every segment
  int scr_cap, cap_ts, n_seg;
};

struct Cta {
  const int4* prog;
  const float* len;
  const int* par;
  const double* Bt;      // [SPAD][SPAD]  Bt[s][a] = B[a][s]
  const float* rate_p;
  const float* rinv_p;
  const double* pi_p;
  const uint8_t* part;
  const double* absorb;  // [S][n_parts][3]  (lam1, sq, r) of the tolerance block per state and class
  const int* tslot;      // [n_nodes] tolerance observation slot
  double* dwell_acc;     // [S]
  unsigned* trans_acc;   // [S*S]
  double* tol_acc;       // [n_parts][4]
  double* sum_acc;       // [8]
};

struct Wp {
  uint8_t* pn;     // [n_nodes]  primary state at the nodes
  uint8_t* pc;     // [n_nodes]  primary jumps on the edge above a node
  uint8_t* sc;     // [n_nodes]  candidate events of the sweep on that edge
  uint8_t* psb;    // [cap_p]
  float* pt;       // [cap_p]
  uint32_t* tn;    // [n_nodes]  tolerance bits at the nodes
  uint8_t* tc;     // [n_nodes][n_parts]  toggles of class c on the edge above a node
  uint8_t* tcc;    // [n_nodes][32]       tolerance candidate events of the sweep, per lane
  float2* tu;      // [scr_cap]  (time, uniform) of the primary candidate events
  double* vec;     // [SPAD]
  double* stk;     // [n_slots][64]
  int* estk;       // [n_slots][32]  power-of-two exponents of the parked tolerance partials (summary)
};

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// state drawn from weights w (lane owns states SP*lane .. SP*lane+SP-1); -1 if all zero
template <int SP>
__device__ __forceinline__ int warp_sample(const double (&w)[SP], uint32_t u, int lane) {
  double loc = w[0];
  if (SP == 2) loc += w[SP - 1];
  double inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  const double tot = __shfl_sync(FULL, inc, 31);
  if (!(tot > 0.0)) return -1;
  const double x = ((double)u + 0.5) * 2.3283064365386963e-10 * tot;
  const double exc = inc - loc;
  int mine = -1;
  if (w[0] > 0.0 && x <= exc + w[0]) mine = SP * lane;
  else if (SP == 2 && w[SP - 1] > 0.0 && x <= inc) mine = SP * lane + 1;
  unsigned hit = __ballot_sync(FULL, mine >= 0);
  if (hit == 0u) {   // rounding at the top of the CDF: the last state with weight
    if (SP == 2 && w[SP - 1] > 0.0) mine = SP * lane + 1;
    else if (w[0] > 0.0) mine = SP * lane;
    hit = __ballot_sync(FULL, mine >= 0);
    const int src = 31 - __clz(hit);
    return __shfl_sync(FULL, mine, src);
  }
  const int src = __ffs(hit) - 1;
  return __shfl_sync(FULL, mine, src);
}

template <int SP>
__device__ __forceinline__ void load_obs_lane(const rt_tmjp_args& A, int slot, int64_t site,
                                              int lane, double (&v)[SP]) {
  if (A.obs_kind == OBS_CODES) {
    const int k = reinterpret_cast<const uint8_t*>(A.obs)[(int64_t)slot * A.obs_stride + site];
#pragma unroll
    for (int j = 0; j < SP; ++j) {
      const int s = SP * lane + j;
      v[j] = (s < A.S && (k == RT_MISSING || k == s)) ? 1.0 : 0.0;
    }
  } else {
    const unsigned long long mk =
        reinterpret_cast<const unsigned long long*>(A.obs)[(int64_t)slot * A.obs_stride + site];
#pragma unroll
    for (int j = 0; j < SP; ++j) {
      const int s = SP * lane + j;
      v[j] = (s < A.S && ((mk >> s) & 1ull)) ? 1.0 : 0.0;
    }
  }
}

// beta <- B beta  (lane owns SP consecutive states; Bt in shared memory, conflict free)
template <int SP>
__device__ __forceinline__ void matvec(const double* Bt, double* vec, double (&beta)[SP], int lane, int S) {
  constexpr int SPAD = SP * 32;
#pragma unroll
  for (int j = 0; j < SP; ++j) vec[SP * lane + j] = beta[j];
  __syncwarp();
  double nb[SP];
#pragma unroll
  for (int j = 0; j < SP; ++j) nb[j] = 0.0;
#pragma unroll 4
  for (int s = 0; s < S; ++s) {
    const double b = vec[s];
#pragma unroll
    for (int j = 0; j < SP; ++j) nb[j] = fma(Bt[s * SPAD + SP * lane + j], b, nb[j]);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < SP; ++j) beta[j] = nb[j];
}

// =====================================================================================
// primary trajectory: one Rao-Teh sweep (or one initial-history attempt)
// =====================================================================================
template <int SP>
__device__ int primary_pass(const rt_tmjp_args& A, const Cta& C, const Wp& W, int lane,
                            int64_t traj, int64_t site, uint32_t sweep, bool init, bool use_tol,
                            double* scr_beta, int scr_cap, int& p_total, bool stats) {
  constexpr int SPAD = SP * 32;
  const int S = A.S;
  const int NP = A.n_parts;
  Philox rng;
  rng.init(A.seed, (uint64_t)(A.traj0 + traj), sweep);
  rng.block(0u, 0u);
  const uint32_t u_root = rng.out[0];
  float hrem = -__logf(RT_U32_TO_UNIT(rng.out[1]));

  bool valid[SP];
  int mypart[SP];
#pragma unroll
  for (int j = 0; j < SP; ++j) {
    const int s = SP * lane + j;
    valid[j] = s < S;
    mypart[j] = (use_tol && valid[j]) ? C.part[s] : 0;
  }
  const bool tl = use_tol && lane < NP;
  const float* my_tt = A.t_time + ((size_t)traj * (size_t)(NP > 0 ? NP : 1) + (size_t)(tl ? lane : 0)) * (size_t)A.cap_t;
  int t_rd = tl ? A.cap_t - (int)A.t_total[(size_t)traj * NP + lane] : 0;
  // the next toggle of this lane's class in pool order, loaded one consumption ahead so that
  // its global-memory latency is never on the (serial) walk
  float pend = (tl && t_rd < A.cap_t) ? my_tt[t_rd] : -1.0f;
  const uint32_t partmask = NP >= 32 ? FULL : ((1u << NP) - 1u);

  double acc[SP];
#pragma unroll
  for (int j = 0; j < SP; ++j) acc[j] = valid[j] ? 1.0 : 0.0;
  int nA = 0;
  int rd = A.cap_p - p_total;
  bool overflow = false, infeasible = false;
  int root_state = 0;

  // =============================== UP ===============================
  for (int ip = 0; ip < A.n_ops; ++ip) {
    const int4 op = C.prog[ip];
    const int code = op.x & 0xff;
    if (code <= OP_MSG_ONES) {
      const int c = op.y;
      double beta[SP];
      if (code == OP_MSG_SLOT) {
#pragma unroll
        for (int j = 0; j < SP; ++j) beta[j] = W.stk[op.z * 64 + SP * lane + j];
      } else if (code == OP_MSG_OBS) {
        load_obs_lane<SP>(A, op.z, site, lane, beta);
      } else {
#pragma unroll
        for (int j = 0; j < SP; ++j) beta[j] = valid[j] ? 1.0 : 0.0;
      }
      const float tc = C.len[c];
      int kA = 0, cur = 0, k_old = 0, next_sb = 0, init_left = 0;
      float next_old = -1.0f, pos = tc;
      uint32_t off_now = 0u, chunk_off = 0u;
      float my_next = -1.0f;
      int my_k = 0;
      if (init) {
        init_left = tc > 0.0f ? A.init_k : 0;   // no events on a zero-length branch (P(0) = I)
      } else {
        cur = W.pn[c];
        k_old = W.pc[c];
        if (k_old > 0) { next_old = W.pt[rd]; next_sb = W.psb[rd]; }
        if (use_tol) {
          off_now = ~W.tn[c] & partmask;
          chunk_off = off_now;
          if (tl) {
            my_k = W.tc[c * NP + lane];
            if (my_k > 0) my_next = pend;
          }
        }
      }
      while (true) {
        float cand;
        bool is_old = false, is_virtual = false;
        if (init) {
          if (init_left == 0) break;
          cand = tc * (float)init_left / (float)(A.init_k + 1);
          --init_left;
        } else {
          float nt = -1.0f;
          if (use_tol) nt = __int_as_float(__reduce_max_sync(FULL, __float_as_int(my_next)));
          const float bound = fmaxf(fmaxf(next_old, nt), 0.0f);
          const float h = C.rate_p[cur] * (pos - bound);
          // a virtual event must fall STRICTLY inside (bound, pos): event times of the
          // primary and of every tolerance class stay pairwise distinct in float32, so the
          // order of a jump and a toggle is the same in every later pass
          bool fire = hrem < h;
          if (fire) {
            cand = pos - hrem * C.rinv_p[cur];
            if (!(cand > bound && cand < pos)) cand = bound + 0.5f * (pos - bound);
            if (!(cand > bound && cand < pos)) { fire = false; hrem = h; }   // sliver below 2 ulp
          }
          if (fire) {
            is_virtual = true;
          } else {
            hrem -= h;
            if (!(bound > 0.0f)) break;
            if (nt >= next_old) {
              // tolerance toggle(s): the classes that change state at nt
              const bool mine = tl && my_next == nt;
              const uint32_t m = __ballot_sync(FULL, mine);
              off_now ^= m;
              chunk_off |= off_now;
              if (mine) {
                ++t_rd;
                --my_k;
                pend = t_rd < A.cap_t ? my_tt[t_rd] : -1.0f;
                my_next = my_k > 0 ? pend : -1.0f;
              }
              pos = nt;
              continue;
            }
            cand = next_old;
            is_old = true;
          }
        }
        if ((kA & 1) == 0) rng.block((uint32_t)ip + 1u, (uint32_t)(kA >> 1));
        const bool odd = (kA & 1) != 0;
        const uint32_t u_draw = odd ? rng.out[2] : rng.out[0];
        const uint32_t u_gap = odd ? rng.out[3] : rng.out[1];
        if (is_virtual) hrem = -__logf(RT_U32_TO_UNIT(u_gap));
        // close the chunk below the event
        if (use_tol && !init) {
#pragma unroll
          for (int j = 0; j < SP; ++j)
            if ((chunk_off >> mypart[j]) & 1u) beta[j] = 0.0;
        }
        if (nA < scr_cap) {
          W.tu[nA] = make_float2(cand, __uint_as_float(u_draw));
#pragma unroll
          for (int j = 0; j < SP; ++j) scr_beta[(size_t)nA * SPAD + SP * lane + j] = beta[j];
        } else {
          overflow = true;
        }
        ++nA;
        ++kA;
        matvec<SP>(C.Bt, W.vec, beta, lane, S);
        chunk_off = off_now;
        pos = cand;
        if (is_old) {
          cur = next_sb;
          ++rd;
          --k_old;
          if (k_old > 0) { next_old = W.pt[rd]; next_sb = W.psb[rd]; }
          else next_old = -1.0f;
        }
      }
      // the piece between the last event and the parent belongs to the parent's chunk
      if (use_tol && !init) {
#pragma unroll
        for (int j = 0; j < SP; ++j)
          if ((chunk_off >> mypart[j]) & 1u) beta[j] = 0.0;
      }
      if (kA > 255) overflow = true;
      W.sc[c] = (uint8_t)(kA > 255 ? 255 : kA);
      if (kA > 0) {
        double mx = beta[0];
        if (SP == 2) mx = fmax(mx, beta[SP - 1]);
        mx = warp_max_d(mx);
        if (mx > 0.0) {
          const double sc = rt_pow2_neg(rt_exponent(mx));
#pragma unroll
          for (int j = 0; j < SP; ++j) beta[j] *= sc;
        }
      }
#pragma unroll
      for (int j = 0; j < SP; ++j) acc[j] *= beta[j];
    } else if (code == OP_APPLY_OBS) {
      double v[SP];
      load_obs_lane<SP>(A, op.z, site, lane, v);
#pragma unroll
      for (int j = 0; j < SP; ++j) acc[j] *= v[j];
    } else {   // OP_STORE / OP_ROOT
      double mx = acc[0];
      if (SP == 2) mx = fmax(mx, acc[SP - 1]);
      mx = warp_max_d(mx);
      if (mx > 0.0) {
        const double sc = rt_pow2_neg(rt_exponent(mx));
#pragma unroll
        for (int j = 0; j < SP; ++j) acc[j] *= sc;
      }
      if (code == OP_STORE) {
#pragma unroll
        for (int j = 0; j < SP; ++j) {
          W.stk[op.z * 64 + SP * lane + j] = acc[j];
          acc[j] = valid[j] ? 1.0 : 0.0;
        }
      } else {
        double w[SP];
#pragma unroll
        for (int j = 0; j < SP; ++j) w[j] = valid[j] ? C.pi_p[SP * lane + j] * acc[j] : 0.0;
        root_state = warp_sample<SP>(w, u_root, lane);
        if (root_state < 0) infeasible = true;
      }
    }
  }
  __syncwarp();
  if (infeasible) return RT_SITE_STRUCTURAL_ZERO;
  if (overflow) return 3;

  // ============================== DOWN ==============================
  W.pn[0] = (uint8_t)root_state;
  int rdA = nA;
  int wr = A.cap_p;
  bool pool_overflow = false;
  double bnext[SP];
#pragma unroll
  for (int j = 0; j < SP; ++j) bnext[j] = nA > 0 ? scr_beta[(size_t)(nA - 1) * SPAD + SP * lane + j] : 0.0;
  for (int ip = A.n_ops - 1; ip >= 0; --ip) {
    const int4 op = C.prog[ip];
    if ((op.x & 0xff) > OP_MSG_ONES) continue;
    const int c = op.y;
    int cur = W.pn[C.par[c]];
    const int kA = W.sc[c];
    const float tc = C.len[c];
    float prev = 0.0f;
    int kept = 0;
    for (int e = 0; e < kA; ++e) {
      --rdA;
      const float2 tu = W.tu[rdA];
      const float tau = tu.x;
      const uint32_t u_draw = __float_as_uint(tu.y);
      double w[SP];
#pragma unroll
      for (int j = 0; j < SP; ++j) {
        const int s = SP * lane + j;
        w[j] = C.Bt[s * SPAD + cur] * bnext[j];
      }
      if (rdA > 0) {   // the record of the next event (whatever edge it is on) while this one is drawn
#pragma unroll
        for (int j = 0; j < SP; ++j) bnext[j] = scr_beta[(size_t)(rdA - 1) * SPAD + SP * lane + j];
      }
      int nxt = warp_sample<SP>(w, u_draw, lane);
      if (nxt < 0) nxt = cur;
      if (stats && lane == 0) atomicAdd(&C.dwell_acc[cur], (double)(tau - prev));
      prev = tau;
      if (nxt != cur) {
        if (stats && lane == 0) atomicAdd(&C.trans_acc[cur * S + nxt], 1u);
        --wr;
        if (wr >= 0) { W.pt[wr] = tau; W.psb[wr] = (uint8_t)cur; }
        else pool_overflow = true;
        ++kept;
        cur = nxt;
      }
    }
    if (stats && lane == 0) atomicAdd(&C.dwell_acc[cur], (double)(tc - prev));
    W.pn[c] = (uint8_t)cur;
    W.pc[c] = (uint8_t)kept;
  }
  __syncwarp();
  if (pool_overflow) { p_total = 0; return 4; }
  p_total = A.cap_p - wr;
  return 0;
}

// =====================================================================================
// tolerance trajectories: lane = class
// =====================================================================================
__device__ int tol_pass(const rt_tmjp_args& A, const Cta& C, const Wp& W, int lane, int64_t traj,
                        int64_t site, uint32_t sweep, bool init, int p_total, float2* scr_tu,
                        double2* scr_b, int cap_ts, bool stats) {
  const int NP = A.n_parts;
  const bool tl = lane < NP;
  Philox rng;
  rng.init(A.seed, (uint64_t)(A.traj0 + traj), sweep);
  const uint32_t sub0 = (uint32_t)(lane + 1) * (uint32_t)(A.n_ops + 1);
  rng.block(sub0, 0u);
  const uint32_t u_root = rng.out[0];
  float hrem = -__logf(RT_U32_TO_UNIT(rng.out[1]));
  const double b01 = A.rate_on / A.omega_t, b00 = 1.0 - b01;
  const double b10 = A.rate_off / A.omega_t, b11 = 1.0 - b10;
  const float r0 = (float)(A.omega_t - A.rate_on), r1 = (float)(A.omega_t - A.rate_off);
  const float ri0 = r0 > 0.0f ? 1.0f / r0 : 0.0f, ri1 = r1 > 0.0f ? 1.0f / r1 : 0.0f;
  const double pi0 = A.rate_off / (A.rate_on + A.rate_off), pi1 = A.rate_on / (A.rate_on + A.rate_off);
  float* my_tt = A.t_time + ((size_t)traj * NP + (tl ? lane : 0)) * (size_t)A.cap_t;
  int rd = (tl && !init) ? A.cap_t - (int)A.t_total[(size_t)traj * NP + lane] : A.cap_t;
  float pend = (tl && rd < A.cap_t) ? my_tt[rd] : -1.0f;   // next old toggle, one consumption ahead
  int prd = A.cap_p - p_total;
  double a0 = 1.0, a1 = 1.0;
  int nA = 0;
  bool overflow = false, infeasible = false;
  int root_state = 0;

  // =============================== UP ===============================
  for (int ip = 0; ip < A.n_ops; ++ip) {
    const int4 op = C.prog[ip];
    const int code = op.x & 0xff;
    if (code <= OP_MSG_ONES) {
      const int c = op.y;
      double be0 = 1.0, be1 = 1.0;
      if (code == OP_MSG_SLOT) {
        be0 = W.stk[(op.z * 2 + 0) * 32 + lane];
        be1 = W.stk[(op.z * 2 + 1) * 32 + lane];
      } else if (tl && A.tol_obs) {
        const int ts = C.tslot[c];
        if (ts >= 0) {
          const int bits = A.tol_obs[((size_t)ts * NP + lane) * (size_t)A.tol_obs_stride + site];
          be0 = (bits & 1) ? 1.0 : 0.0;
          be1 = (bits & 2) ? 1.0 : 0.0;
        }
      }
      const float tc = C.len[c];
      const int pk0 = W.pc[c];
      int kA = 0;
      if (tl) {
        int pstate = W.pn[c];
        int pk = pk0;
        int pr = prd;
        float next_p = pk > 0 ? W.pt[pr] : -1.0f;
        int next_psb = pk > 0 ? W.psb[pr] : 0;
        bool need_on = C.part[pstate] == lane;
        int cur_t = (W.tn[c] >> lane) & 1;
        int k_old = init ? 0 : W.tc[c * NP + lane];
        float next_old = k_old > 0 ? pend : -1.0f;
        float pos = tc;
        bool placed = false;
        while (true) {
          const float bound = fmaxf(fmaxf(next_old, next_p), 0.0f);
          float cand;
          bool is_old = false, is_virtual = false, is_event = true;
          if (init) {
            if (placed || !(pos > bound)) {     // (no event inside a zero-length segment)
              is_event = false;
              cand = bound;
            } else {
              cand = 0.0f;   // set below from the event's uniform
              placed = true;
            }
          } else {
            const float h = (cur_t ? r1 : r0) * (pos - bound);
            bool fire = hrem < h;
            if (fire) {   // strictly inside (bound, pos), see primary_pass
              cand = pos - hrem * (cur_t ? ri1 : ri0);
              if (!(cand > bound && cand < pos)) cand = bound + 0.5f * (pos - bound);
              if (!(cand > bound && cand < pos)) { fire = false; hrem = h; }
            }
            if (fire) {
              is_virtual = true;
            } else {
              hrem -= h;
              if (next_p >= next_old) { is_event = false; cand = bound; }
              else { cand = next_old; is_old = true; }
            }
          }
          if (!is_event) {
            if (!(bound > 0.0f)) break;
            // a primary jump: the class may be required ON above it
            pstate = next_psb;
            need_on = need_on || (C.part[pstate] == lane);
            ++pr;
            --pk;
            next_p = pk > 0 ? W.pt[pr] : -1.0f;
            next_psb = pk > 0 ? W.psb[pr] : 0;
            pos = bound;
            placed = false;
            continue;
          }
          if ((kA & 1) == 0) rng.block(sub0 + (uint32_t)ip + 1u, (uint32_t)(kA >> 1));
          const bool odd = (kA & 1) != 0;
          const uint32_t u_draw = odd ? rng.out[2] : rng.out[0];
          const uint32_t u_gap = odd ? rng.out[3] : rng.out[1];
          if (is_virtual) hrem = -__logf(RT_U32_TO_UNIT(u_gap));
          if (init) {
            cand = bound + (pos - bound) * (((float)(u_gap >> 8) + 0.5f) * 5.9604644775390625e-8f);
            if (!(cand > bound) || !(cand < pos)) cand = bound + 0.5f * (pos - bound);
          }
          if (need_on) be0 = 0.0;
          if (nA < cap_ts) {
            scr_tu[(size_t)nA * 32 + lane] = make_float2(cand, __uint_as_float(u_draw));
            scr_b[(size_t)nA * 32 + lane] = make_double2(be0, be1);
          } else {
            overflow = true;
          }
          ++nA;
          ++kA;
          {
            const double n0 = b00 * be0 + b01 * be1;
            const double n1 = b10 * be0 + b11 * be1;
            be0 = n0;
            be1 = n1;
          }
          need_on = C.part[pstate] == lane;
          pos = cand;
          if (is_old) {
            cur_t ^= 1;
            ++rd;
            --k_old;
            pend = rd < A.cap_t ? my_tt[rd] : -1.0f;
            next_old = k_old > 0 ? pend : -1.0f;
          }
        }
        if (need_on) be0 = 0.0;
        if (kA > 255) overflow = true;
        W.tcc[c * 32 + lane] = (uint8_t)(kA > 255 ? 255 : kA);
        if (kA > 0) {
          const double mx = fmax(be0, be1);
          if (mx > 0.0) {
            const double sc = rt_pow2_neg(rt_exponent(mx));
            be0 *= sc;
            be1 *= sc;
          }
        }
      }
      prd += pk0;
      a0 *= be0;
      a1 *= be1;
    } else if (code >= OP_STORE) {
      const int v = op.y;
      if (tl && A.tol_obs) {
        const int ts = C.tslot[v];
        if (ts >= 0) {
          const int bits = A.tol_obs[((size_t)ts * NP + lane) * (size_t)A.tol_obs_stride + site];
          if (!(bits & 1)) a0 = 0.0;
          if (!(bits & 2)) a1 = 0.0;
        }
      }
      const double mx = fmax(a0, a1);
      if (mx > 0.0) {
        const double sc = rt_pow2_neg(rt_exponent(mx));
        a0 *= sc;
        a1 *= sc;
      }
      if (code == OP_STORE) {
        W.stk[(op.z * 2 + 0) * 32 + lane] = a0;
        W.stk[(op.z * 2 + 1) * 32 + lane] = a1;
        a0 = 1.0;
        a1 = 1.0;
      } else {
        const double w0 = pi0 * a0, w1 = pi1 * a1;
        const double tot = w0 + w1;
        if (!(tot > 0.0)) infeasible = true;
        const double x = ((double)u_root + 0.5) * 2.3283064365386963e-10 * tot;
        root_state = (w0 > 0.0 && (x <= w0 || !(w1 > 0.0))) ? 0 : 1;
      }
    }
  }
  const unsigned bad_inf = __ballot_sync(FULL, tl && infeasible);
  const unsigned bad_ovf = __ballot_sync(FULL, tl && overflow);
  if (bad_inf) return 6;   // no feasible tolerance history
  if (bad_ovf) return 3;

  // ============================== DOWN ==============================
  W.tn[0] = __ballot_sync(FULL, tl && root_state == 1);
  int rdA = nA;
  int wr = A.cap_t;
  bool pool_overflow = false;
  double dwell_on = 0.0, gains = 0.0, losses = 0.0;
  float2 tu_next = make_float2(0.0f, 0.0f);
  double2 b_next = make_double2(0.0, 0.0);
  if (tl && nA > 0) {
    tu_next = scr_tu[(size_t)(nA - 1) * 32 + lane];
    b_next = scr_b[(size_t)(nA - 1) * 32 + lane];
  }
  for (int ip = A.n_ops - 1; ip >= 0; --ip) {
    const int4 op = C.prog[ip];
    if ((op.x & 0xff) > OP_MSG_ONES) continue;
    const int c = op.y;
    int cur = (W.tn[C.par[c]] >> lane) & 1;
    if (tl) {
      const int kA = W.tcc[c * 32 + lane];
      const float tc = C.len[c];
      float prev = 0.0f;
      int kept = 0;
      for (int e = 0; e < kA; ++e) {
        --rdA;
        const float2 tu = tu_next;
        const double2 b = b_next;
        if (rdA > 0) {   // next record of this lane while the current event is drawn
          tu_next = scr_tu[(size_t)(rdA - 1) * 32 + lane];
          b_next = scr_b[(size_t)(rdA - 1) * 32 + lane];
        }
        const float tau = tu.x;
        const uint32_t u_draw = __float_as_uint(tu.y);
        const double w0 = (cur ? b10 : b00) * b.x, w1 = (cur ? b11 : b01) * b.y;
        const double tot = w0 + w1;
        const double x = ((double)u_draw + 0.5) * 2.3283064365386963e-10 * tot;
        int nxt = (w0 > 0.0 && (x <= w0 || !(w1 > 0.0))) ? 0 : 1;
        if (!(tot > 0.0)) nxt = cur;
        if (cur) dwell_on += (double)(tau - prev);
        prev = tau;
        if (nxt != cur) {
          if (nxt) gains += 1.0; else losses += 1.0;
          --wr;
          if (wr >= 0) my_tt[wr] = tau; else pool_overflow = true;
          ++kept;
          cur = nxt;
        }
      }
      if (cur) dwell_on += (double)(tc - prev);
      W.tc[c * NP + lane] = (uint8_t)kept;
    }
    W.tn[c] = __ballot_sync(FULL, tl && cur == 1);
  }
  const unsigned bad_pool = __ballot_sync(FULL, tl && pool_overflow);
  if (tl) A.t_total[(size_t)traj * NP + lane] = (uint8_t)(pool_overflow ? 0 : A.cap_t - wr);
  if (bad_pool) return 4;
  if (stats && tl) {
    double* acc = C.tol_acc + lane * 4;
    atomicAdd(&acc[0], (double)root_state);
    atomicAdd(&acc[1], dwell_on);
    atomicAdd(&acc[2], gains);
    atomicAdd(&acc[3], losses);
  }
  return 0;
}

// =====================================================================================
// Rao-Blackwellised tolerance summary of the primary trajectory: lane = class
// =====================================================================================
__constant__ double kInvFact1[18] = {   // 1/(k+1)!
    1.0, 0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333,
    0.001388888888888889, 0.0001984126984126984, 2.48015873015873e-05, 2.7557319223985893e-06,
    2.755731922398589e-07, 2.505210838544172e-08, 2.08767569878681e-09, 1.6059043836821613e-10,
    1.1470745597729725e-11, 7.647163731819816e-13, 4.779477332387385e-14, 2.8114572543455206e-15,
    1.5619206968586225e-16};

// phi1 = (e^x - 1)/x, phi2 = (e^x - 1 - x)/x^2, g3 = (x e^x - 2(e^x - 1) + x)/x^3 for x <= 0:
// p1 = sum x^k/(k+1)!, p2 = sum x^k/(k+2)!, g3 = sum (k+1) x^k/(k+3)! below |x| = 1/2.
struct Seg {          // A = [[-a, a], [w, -w-r]] = top-left block of Q3 (_linalg.py:14-29)
  double n00, n01, n10, n11;   // N = A - lam1 I
  double e1t, psit, I1, I2, t;
  double p00, p01, p10, p11;   // exp(tA)
};

// Per (primary state, class) constants of A (computed once per CTA): lam1 (the eigenvalue
// nearer 0), sq = lam1 - lam2 >= 0, r (absorption rate).  lam1 lam2 = det A = a r.
__device__ __forceinline__ void seg_constants(double a, double w, double r, double& lam1, double& sq) {
  const double tr = -(a + w + r);
  const double disc = (a - r) * (a - r) + w * w + 2.0 * w * (a + r);
  sq = sqrt(disc);
  const double lam2 = 0.5 * (tr - sq);
  lam1 = lam2 < 0.0 ? (a * r) / lam2 : 0.0;
}

// exp(tA) = e^{lam1 t} (I + psi(t) N), psi(t) = t phi1(x), x = -sq t.  (Keeping e^{lam1 t} and
// expm1(x) from the upward pass for the downward pass was measured slower than recomputing:
// most segments are short, |x| < 1/2, where the series needs no expm1 at all.)
__device__ __forceinline__ void seg_setup(double a, double w, double r, double lam1, double sq,
                                          double t, bool frechet, Seg& s) {
  const double x = -sq * t;
  s.t = t;
  s.n00 = -a - lam1; s.n01 = a; s.n10 = w; s.n11 = -w - r - lam1;
  s.e1t = exp(lam1 * t);
  double p1;
  if (x > -0.5) {
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 15; k >= 0; --k) {
      s1 = fma(s1, x, kInvFact1[k]);
      if (frechet) {
        s2 = fma(s2, x, kInvFact1[k + 1]);
        s3 = fma(s3, x, kInvFact1[k + 2] * (double)(k + 1));
      }
    }
    p1 = s1;
    if (frechet) { s.I1 = t * t * s2; s.I2 = t * t * t * s3; }
  } else {
    const double em1 = expm1(x);
    const double ix = 1.0 / x;
    p1 = em1 * ix;
    if (frechet) {
      s.I1 = t * t * ((em1 - x) * ix * ix);
      s.I2 = t * t * t * ((x * (em1 + 2.0) - 2.0 * em1) * ix * ix * ix);
    }
  }
  s.psit = t * p1;
  s.p00 = s.e1t * (1.0 + s.psit * s.n00);
  s.p01 = s.e1t * s.psit * s.n01;
  s.p10 = s.e1t * s.psit * s.n10;
  s.p11 = s.e1t * (1.0 + s.psit * s.n11);
  if (s.p00 < 0.0) s.p00 = 0.0;      // tiny negative round-off on the diagonal
  if (s.p11 < 0.0) s.p11 = 0.0;
}

// out8 (valid on every lane after the call): initial_on, initial_off, dwell_on, dwell_off,
// nabsorptions, ngains, nlosses  (raoteh/sampler/_tmjp_dense.py:852-855), and [7] = the
// log-likelihood of the primary trajectory under the compound process with the tolerance
// histories integrated out (get_tolerance_process_log_likelihood, _tmjp_dense.py:407-505):
// log pi(root state) + sum over jumps log Q[s,s'] + sum over classes log L_c, where class c
// starts from (0, 1) if the root's primary state belongs to it and from tolerance_distn else.
template <int SP>
__device__ int summary_pass(const rt_tmjp_args& A, const Cta& C, const Wp& W, int lane, int64_t site,
                            int p_total, double2* scr_seg, int n_seg_cap, double total_len,
                            double (&out7)[8], const double* pt64) {
  const double* len64 = pt64 ? A.length : nullptr;   // fp64 branch lengths go with fp64 jump times
  const int NP = A.n_parts;
  const bool tl = lane < NP;
  const double a = A.rate_on;
  const double pi0 = A.rate_off / (A.rate_on + A.rate_off), pi1 = A.rate_on / (A.rate_on + A.rate_off);
  double a0 = 1.0, a1 = 1.0;
  int esum = 0;                    // exponent of the partial being built (true = value * 2^esum)
  int nseg = 0;
  int prd = A.cap_p - p_total;
  bool bad = false, overflow = false;
  double d0 = 0.0, d1 = 0.0;
  double class_ll = 0.0;           // log L_c of this lane's class
  double jump_ll = 0.0;            // sum over primary jumps of log Q[s, s'] (same on every lane)

  // =============================== UP ===============================
  for (int ip = 0; ip < A.n_ops; ++ip) {
    const int4 op = C.prog[ip];
    const int code = op.x & 0xff;
    if (code <= OP_MSG_ONES) {
      const int c = op.y;
      double be0 = 1.0, be1 = 1.0;
      if (code == OP_MSG_SLOT) {
        be0 = W.stk[(op.z * 2 + 0) * 32 + lane];
        be1 = W.stk[(op.z * 2 + 1) * 32 + lane];
        esum += W.estk[op.z * 32 + lane];
      } else if (tl && A.tol_obs) {
        const int ts = C.tslot[c];
        if (ts >= 0) {
          const int bits = A.tol_obs[((size_t)ts * NP + lane) * (size_t)A.tol_obs_stride + site];
          be0 = (bits & 1) ? 1.0 : 0.0;
          be1 = (bits & 2) ? 1.0 : 0.0;
        }
      }
      const int pk0 = W.pc[c];
      if (tl) {
        int pstate = W.pn[c];
        int pk = pk0, pr = prd;
        double pos = len64 ? len64[c] : (double)C.len[c];
        while (true) {
          const double bound = pk > 0 ? (pt64 ? pt64[pr] : (double)W.pt[pr]) : 0.0;
          const bool same = C.part[pstate] == lane;
          const double w = same ? 0.0 : A.rate_off;
          const double* ct = C.absorb + (size_t)(pstate * NP + lane) * 3;   // (lam1, sq, r)
          const double t = pos - bound;
          if (same) be0 = 0.0;
          if (nseg < n_seg_cap) scr_seg[(size_t)nseg * 32 + lane] = make_double2(be0, be1);
          else overflow = true;
          ++nseg;
          Seg s;
          seg_setup(a, w, ct[2], ct[0], ct[1], t, false, s);
          const double n0 = s.p00 * be0 + s.p01 * be1;
          const double n1 = s.p10 * be0 + s.p11 * be1;
          be0 = same ? 0.0 : n0;
          be1 = n1;
          if (pk == 0) break;
          pstate = W.psb[pr];
          ++pr;
          --pk;
          pos = bound;
        }
      }
      if (lane == 0) {   // the primary jumps of this edge: rate Q[parent side][child side]
        int below = W.pn[c];
        for (int i = 0; i < pk0; ++i) {
          const int above = W.psb[prd + i];
          jump_ll += log(C.Bt[below * (SP * 32) + above] * A.omega_p);
          below = above;
        }
      }
      prd += pk0;
      a0 *= be0;
      a1 *= be1;
    } else if (code >= OP_STORE) {
      const int v = op.y;
      if (tl && A.tol_obs) {
        const int ts = C.tslot[v];
        if (ts >= 0) {
          const int bits = A.tol_obs[((size_t)ts * NP + lane) * (size_t)A.tol_obs_stride + site];
          if (!(bits & 1)) a0 = 0.0;
          if (!(bits & 2)) a1 = 0.0;
        }
      }
      const double mx = fmax(a0, a1);
      if (mx > 0.0) {
        const int e = rt_exponent(mx);
        const double sc = rt_pow2_neg(e);
        a0 *= sc;
        a1 *= sc;
        esum += e;
      }
      if (code == OP_STORE) {
        W.stk[(op.z * 2 + 0) * 32 + lane] = a0;
        W.stk[(op.z * 2 + 1) * 32 + lane] = a1;
        W.estk[op.z * 32 + lane] = esum;
        a0 = 1.0;
        a1 = 1.0;
        esum = 0;
      } else {
        const double w0 = pi0 * a0, w1 = pi1 * a1;
        const double tot = w0 + w1;
        if (!(tot > 0.0)) bad = true;
        else { d0 = w0 / tot; d1 = w1 / tot; }
        // likelihood prior of the class: (0, 1) when the root's primary state belongs to it
        const bool own = C.part[W.pn[0]] == lane;
        const double lk = own ? a1 : tot;
        class_ll = lk > 0.0 ? log(lk) + (double)esum * RT_LN2 : -INFINITY;
      }
    }
  }
  const unsigned any_bad = __ballot_sync(FULL, tl && bad);
  const unsigned any_ovf = __ballot_sync(FULL, tl && overflow);
  if (any_ovf) return 3;
  if (any_bad) return RT_SITE_NUMERICAL_ZERO;

  // ============================== DOWN ==============================
  double init_on = d1, dwell_on = 0.0, nabs = 0.0, gains = 0.0, losses = 0.0;
  double dc0 = d0, dc1 = d1;       // marginal of the node whose child edges are being walked
  int sg = nseg;
  int pr = A.cap_p - 1;            // the jump list read backwards = down order
  double2 l_next = make_double2(0.0, 0.0);
  if (tl && nseg > 0) l_next = scr_seg[(size_t)(nseg - 1) * 32 + lane];
  for (int ip = A.n_ops - 1; ip >= 0; --ip) {
    const int4 op = C.prog[ip];
    const int code = op.x & 0xff;
    if (code == OP_STORE) {
      dc0 = W.stk[(op.z * 2 + 0) * 32 + lane];
      dc1 = W.stk[(op.z * 2 + 1) * 32 + lane];
      continue;
    }
    if (code > OP_MSG_ONES) continue;
    const int c = op.y;
    const int k = W.pc[c];
    double t0 = dc0, t1 = dc1;
    if (tl) {
      double tprev = 0.0;
      for (int i = 0; i <= k; ++i) {
        int st;
        double tend;
        if (i < k) { st = W.psb[pr - i]; tend = pt64 ? pt64[pr - i] : (double)W.pt[pr - i]; }
        else { st = W.pn[c]; tend = len64 ? len64[c] : (double)C.len[c]; }
        --sg;
        const double2 l = l_next;
        if (sg > 0) l_next = scr_seg[(size_t)(sg - 1) * 32 + lane];   // next record meanwhile
        const bool same = C.part[st] == lane;
        const double w = same ? 0.0 : A.rate_off;
        const double* ct = C.absorb + (size_t)(st * NP + lane) * 3;
        const double r = ct[2];
        Seg s;
        seg_setup(a, w, r, ct[0], ct[1], tend - tprev, true, s);
        const double m0 = s.p00 * l.x + s.p01 * l.y;
        const double m1 = s.p10 * l.x + s.p11 * l.y;
        const double g0 = (t0 > 0.0 && m0 > 0.0) ? t0 / m0 : 0.0;
        const double g1 = (t1 > 0.0 && m1 > 0.0) ? t1 / m1 : 0.0;
        // M^{cd} = sum_ab G_a L_b int_0^t exp(sA)[a,c] exp((t-s)A)[d,b] ds
        const double gN0 = g0 * s.n00 + g1 * s.n10, gN1 = g0 * s.n01 + g1 * s.n11;
        const double Nl0 = s.n00 * l.x + s.n01 * l.y, Nl1 = s.n10 * l.x + s.n11 * l.y;
        const double M11 = s.e1t * (g1 * l.y * s.t + (gN1 * l.y + g1 * Nl1) * s.I1 + gN1 * Nl1 * s.I2);
        const double M01 = s.e1t * (g0 * l.y * s.t + (gN0 * l.y + g0 * Nl1) * s.I1 + gN0 * Nl1 * s.I2);
        const double M10 = s.e1t * (g1 * l.x * s.t + (gN1 * l.x + g1 * Nl0) * s.I1 + gN1 * Nl0 * s.I2);
        dwell_on += M11;
        nabs += r * M11;
        gains += a * M01;
        losses += w * M10;
        const double q0 = l.x * (g0 * s.p00 + g1 * s.p10);
        const double q1 = l.y * (g0 * s.p01 + g1 * s.p11);
        t0 = q0;
        t1 = q1;
        tprev = tend;
      }
    }
    pr -= k;
    if (code == OP_MSG_SLOT) {
      W.stk[(op.z * 2 + 0) * 32 + lane] = t0;
      W.stk[(op.z * 2 + 1) * 32 + lane] = t1;
    }
  }
  if (!tl) { init_on = 0.0; dwell_on = 0.0; nabs = 0.0; gains = 0.0; losses = 0.0; }
  init_on = rt_warp_sum(init_on);
  dwell_on = rt_warp_sum(dwell_on);
  nabs = rt_warp_sum(nabs);
  gains = rt_warp_sum(gains);
  losses = rt_warp_sum(losses);
  out7[0] = init_on;
  out7[1] = (double)NP - init_on;
  out7[2] = dwell_on;
  out7[3] = total_len * (double)NP - dwell_on;
  out7[4] = nabs;
  out7[5] = gains;
  out7[6] = losses;
  if (!tl) class_ll = 0.0;
  class_ll = rt_warp_sum(class_ll);
  jump_ll = __shfl_sync(FULL, jump_ll, 0);
  const double prior = C.pi_p[W.pn[0]];
  out7[7] = (prior > 0.0 ? log(prior) : -INFINITY) + jump_ll + class_ll;
  return 0;
}

// log-likelihood of the primary trajectory under the MJP (B, omega_p, pi_p) itself
// (_mjp.get_trajectory_log_likelihood, raoteh/sampler/_mjp.py:186-250):
// log pi(root) - sum_s dwell_s q_s + sum over jumps log Q[s, s'].  Lanes take 32 edges at a time;
// the offset of an edge's jumps in the pool is a warp prefix sum over the program order.
template <int SP>
__device__ double trajectory_loglik(const rt_tmjp_args& A, const Cta& C, const Wp& W, int lane,
                                    int p_total, bool stats, const double* pt64) {
  const double* len64 = pt64 ? A.length : nullptr;
  constexpr int SPAD = SP * 32;
  double ll = 0.0;
  int base = A.cap_p - p_total;
  for (int ip0 = 0; ip0 < A.n_ops; ip0 += 32) {
    const int ip = ip0 + lane;
    int c = -1, k = 0;
    if (ip < A.n_ops) {
      const int4 op = C.prog[ip];
      if ((op.x & 0xff) <= OP_MSG_ONES) { c = op.y; k = W.pc[c]; }
    }
    int inc = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += t;
    }
    const int start = base + inc - k;
    base += __shfl_sync(FULL, inc, 31);
    if (c >= 0) {
      // jumps of the edge, child end first: (time from the parent end, parent-side state)
      int below = W.pn[c];
      double pos = len64 ? len64[c] : (double)C.len[c];
      for (int i = 0; i < k; ++i) {
        const double tau = pt64 ? pt64[start + i] : (double)W.pt[start + i];
        const int above = W.psb[start + i];
        ll -= (pos - tau) * (A.omega_p - A.rate_p[below]);
        ll += log(C.Bt[below * SPAD + above] * A.omega_p);
        if (stats) {   // _mjp_dense.get_history_statistics of the current history (:150)
          atomicAdd(&C.dwell_acc[below], pos - tau);
          atomicAdd(&C.trans_acc[above * A.S + below], 1u);
        }
        below = above;
        pos = tau;
      }
      ll -= pos * (A.omega_p - A.rate_p[below]);
      if (stats) atomicAdd(&C.dwell_acc[below], pos);
    }
  }
  ll = rt_warp_sum(ll);
  const double prior = C.pi_p[W.pn[0]];
  return ll + (prior > 0.0 ? log(prior) : -INFINITY);
}

// =====================================================================================
struct Layout {     // dynamic shared memory carve-up (bytes)
  size_t prog, len, par, Bt, rate_p, rinv_p, pi_p, part, absorb, tslot, dwell, trans, tol, sum;
  size_t warp0, w_pn, w_pc, w_sc, w_psb, w_pt, w_tn, w_tc, w_tcc, w_tu, w_vec, w_stk, w_estk, warp_bytes, total;
};

__host__ __device__ inline size_t al(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline Layout make_layout(int S, int SPAD, int n_parts, int n_nodes, int n_ops,
                                              int n_slots, int cap_p, int scr_cap, int kWarps) {
  Layout L;
  size_t o = 0;
  L.prog = o; o += sizeof(int4) * (size_t)n_ops;
  L.Bt = o = al(o, 16); o += sizeof(double) * (size_t)SPAD * SPAD;
  L.pi_p = o; o += sizeof(double) * (size_t)SPAD;
  L.absorb = o; o += sizeof(double) * 3 * (size_t)S * (size_t)(n_parts > 0 ? n_parts : 1);
  L.dwell = o; o += sizeof(double) * (size_t)S;
  L.tol = o; o += sizeof(double) * 4 * (size_t)(n_parts > 0 ? n_parts : 1);
  L.sum = o; o += sizeof(double) * 8;
  L.trans = o; o += sizeof(unsigned) * (size_t)S * S;
  L.len = o; o += sizeof(float) * (size_t)n_nodes;
  L.par = o; o += sizeof(int) * (size_t)n_nodes;
  L.tslot = o; o += sizeof(int) * (size_t)n_nodes;
  L.rate_p = o; o += sizeof(float) * (size_t)SPAD;
  L.rinv_p = o; o += sizeof(float) * (size_t)SPAD;
  L.part = o; o += al((size_t)SPAD, 4);
  L.warp0 = o = al(o, 16);
  size_t w = 0;
  L.w_vec = w; w += sizeof(double) * 64;
  L.w_stk = w; w += sizeof(double) * 64 * (size_t)n_slots;
  L.w_estk = w; w += n_parts > 0 ? sizeof(int) * 32 * (size_t)n_slots : 0;
  L.w_tu = w; w += sizeof(float2) * (size_t)scr_cap;
  L.w_pt = w; w += sizeof(float) * (size_t)cap_p;
  L.w_tn = w; w += sizeof(uint32_t) * (size_t)n_nodes;
  L.w_pn = w; w += (size_t)n_nodes;
  L.w_pc = w; w += (size_t)n_nodes;
  L.w_sc = w; w += (size_t)n_nodes;
  L.w_psb = w; w += (size_t)cap_p;
  L.w_tc = w; w += (size_t)n_nodes * (size_t)n_parts;
  L.w_tcc = w; w += n_parts > 0 ? (size_t)n_nodes * 32 : 0;
  L.warp_bytes = al(w, 16);
  L.total = L.warp0 + L.warp_bytes * kWarps;
  return L;
}

template <int SP, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, kMaxWarps / kWarps)
tmjp_kernel(rt_tmjp_args A, Scratch X) {
  constexpr int SPAD = SP * 32;
  constexpr int kThreads = kWarps * 32;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double total_len_s;
  const int S = A.S, NP = A.n_parts;
  const int scr_cap = X.scr_cap;
  const Layout L = make_layout(S, SPAD, NP, A.n_nodes, A.n_ops, A.n_slots, A.cap_p, scr_cap, kWarps);
  int4* prog_s = reinterpret_cast<int4*>(smem + L.prog);
  double* Bt_s = reinterpret_cast<double*>(smem + L.Bt);
  double* pi_s = reinterpret_cast<double*>(smem + L.pi_p);
  double* absorb_s = reinterpret_cast<double*>(smem + L.absorb);
  double* dwell_s = reinterpret_cast<double*>(smem + L.dwell);
  double* tol_s = reinterpret_cast<double*>(smem + L.tol);
  double* sum_s = reinterpret_cast<double*>(smem + L.sum);
  unsigned* trans_s = reinterpret_cast<unsigned*>(smem + L.trans);
  float* len_s = reinterpret_cast<float*>(smem + L.len);
  int* par_s = reinterpret_cast<int*>(smem + L.par);
  int* tslot_s = reinterpret_cast<int*>(smem + L.tslot);
  float* rate_s = reinterpret_cast<float*>(smem + L.rate_p);
  float* rinv_s = reinterpret_cast<float*>(smem + L.rinv_p);
  uint8_t* part_s = smem + L.part;

  const int tid = threadIdx.x;
  for (int i = tid; i < A.n_ops; i += kThreads) prog_s[i] = reinterpret_cast<const int4*>(A.program)[i];
  for (int i = tid; i < SPAD * SPAD; i += kThreads) {
    const int s = i / SPAD, a2 = i % SPAD;    // Bt[s][a] = B[a][s]
    Bt_s[i] = (s < S && a2 < S) ? A.B[a2 * S + s] : 0.0;
  }
  for (int i = tid; i < SPAD; i += kThreads) {
    pi_s[i] = i < S ? (A.pi_p ? A.pi_p[i] : 1.0) : 0.0;
    const double r = i < S ? A.rate_p[i] : 0.0;
    rate_s[i] = (float)r;
    rinv_s[i] = r > 0.0 ? (float)(1.0 / r) : 0.0f;
    part_s[i] = (i < S && A.part) ? A.part[i] : 255;
  }
  for (int i = tid; i < S * NP; i += kThreads) {
    const double r = A.absorb ? A.absorb[i] : 0.0;
    const bool same = A.part && A.part[i / NP] == (i % NP);
    double lam1, sq;
    seg_constants(A.rate_on, same ? 0.0 : A.rate_off, r, lam1, sq);
    absorb_s[3 * i] = lam1;
    absorb_s[3 * i + 1] = sq;
    absorb_s[3 * i + 2] = r;
  }
  for (int i = tid; i < A.n_nodes; i += kThreads) {
    len_s[i] = (float)A.length[i];
    par_s[i] = A.parent[i];
    tslot_s[i] = A.tol_obs_slot ? A.tol_obs_slot[i] : -1;
  }
  for (int i = tid; i < S; i += kThreads) dwell_s[i] = 0.0;
  for (int i = tid; i < S * S; i += kThreads) trans_s[i] = 0u;
  for (int i = tid; i < 4 * NP; i += kThreads) tol_s[i] = 0.0;
  if (tid < 8) sum_s[tid] = 0.0;
  if (tid == 0) {   // tree length: dwell_off = total * n_parts - dwell_on (_tmjp_dense.py:850)
    double t = 0.0;
    for (int i = 1; i < A.n_nodes; ++i) t += A.length[i];
    total_len_s = t;
  }
  __syncthreads();
  const double total_len = total_len_s;

  Cta C;
  C.prog = prog_s; C.len = len_s; C.par = par_s; C.Bt = Bt_s; C.rate_p = rate_s; C.rinv_p = rinv_s;
  C.pi_p = pi_s; C.part = part_s; C.absorb = absorb_s; C.tslot = tslot_s;
  C.dwell_acc = dwell_s; C.trans_acc = trans_s; C.tol_acc = tol_s; C.sum_acc = sum_s;

  const int warp = tid >> 5, lane = tid & 31;
  unsigned char* wb = smem + L.warp0 + L.warp_bytes * warp;
  Wp W;
  W.vec = reinterpret_cast<double*>(wb + L.w_vec);
  W.stk = reinterpret_cast<double*>(wb + L.w_stk);
  W.estk = reinterpret_cast<int*>(wb + L.w_estk);
  W.tu = reinterpret_cast<float2*>(wb + L.w_tu);
  W.pt = reinterpret_cast<float*>(wb + L.w_pt);
  W.tn = reinterpret_cast<uint32_t*>(wb + L.w_tn);
  W.pn = wb + L.w_pn; W.pc = wb + L.w_pc; W.sc = wb + L.w_sc; W.psb = wb + L.w_psb;
  W.tc = wb + L.w_tc; W.tcc = wb + L.w_tcc;

  const int64_t gw = (int64_t)blockIdx.x * kWarps + warp;
  const int64_t nw = (int64_t)gridDim.x * kWarps;
  double* scr_beta = X.p_beta + (size_t)gw * (size_t)scr_cap * SPAD;
  float2* scr_tu = X.t_tu ? X.t_tu + (size_t)gw * (size_t)X.cap_ts * 32 : nullptr;
  double2* scr_tb = X.t_beta ? X.t_beta + (size_t)gw * (size_t)X.cap_ts * 32 : nullptr;
  double2* scr_seg = X.seg ? X.seg + (size_t)gw * (size_t)X.n_seg * 32 : nullptr;
  const bool stats_p = (A.flags & RT_TMJP_F_STATS_PRIMARY) != 0;
  const bool stats_t = (A.flags & RT_TMJP_F_STATS_TOLERANCE) != 0;

  for (int64_t traj = gw; traj < A.n_traj; traj += nw) {
    if (A.status[traj] != RT_SITE_OK) continue;
    const int64_t site = (A.traj0 + traj) % A.n_sites;
    // ---- trajectory -> shared memory
    for (int v = lane; v < A.n_nodes; v += 32) {
      W.pn[v] = A.p_node[traj * A.pn_traj_stride + (int64_t)v * A.pn_node_stride];
      W.pc[v] = A.p_cnt[traj * A.pn_traj_stride + (int64_t)v * A.pn_node_stride];
      W.tn[v] = (NP > 0 && A.t_node) ? A.t_node[(size_t)traj * A.n_nodes + v] : 0u;
    }
    for (int i = lane; i < A.cap_p; i += 32) {
      W.pt[i] = A.p_time[(size_t)traj * A.cap_p + i];
      W.psb[i] = A.p_sb[(size_t)traj * A.cap_p + i];
    }
    if (NP > 0)
      for (int i = lane; i < A.n_nodes * NP; i += 32) W.tc[i] = A.t_cnt[(size_t)traj * A.n_nodes * NP + i];
    int p_total = A.p_total[traj];
    // optional fp64 copy of the jump times (caller-loaded trajectories): used by the summary and
    // log-likelihood modes instead of the float32 sampler state
    const double* pt64 = (A.p_time64 && A.mode >= RT_TMJP_SUMMARY) ? A.p_time64 + (size_t)traj * A.cap_p : nullptr;
    __syncwarp();
    int st = 0;
    bool primary_dirty = false, tol_dirty = false;
    double out7[8];
    bool have7 = false;
    if (A.mode == RT_TMJP_INIT_PRIMARY) {
      st = primary_pass<SP>(A, C, W, lane, traj, site, (uint32_t)A.sweep0, true, false, scr_beta,
                            scr_cap, p_total, false);
      primary_dirty = st == 0 || st == 4;
    } else if (A.mode == RT_TMJP_INIT_TOLERANCE) {
      st = tol_pass(A, C, W, lane, traj, site, (uint32_t)A.sweep0, true, p_total, scr_tu, scr_tb,
                    X.cap_ts, false);
      tol_dirty = st == 0 || st == 4;
    } else if (A.mode == RT_TMJP_SWEEP) {
      for (int sw = 0; sw < A.n_sweeps; ++sw) {
        const uint32_t sweep = (uint32_t)(A.sweep0 + sw);
        if (!(A.flags & RT_TMJP_F_SKIP_PRIMARY)) {
          st = primary_pass<SP>(A, C, W, lane, traj, site, sweep, false, NP > 0, scr_beta, scr_cap,
                                p_total, stats_p);
          if (st == 0 || st == 4) primary_dirty = true;
          if (st) break;
        }
        if (NP > 0 && !(A.flags & RT_TMJP_F_SKIP_TOLERANCE)) {
          st = tol_pass(A, C, W, lane, traj, site, sweep, false, p_total, scr_tu, scr_tb,
                        X.cap_ts, stats_t);
          if (st == 0 || st == 4) tol_dirty = true;
          if (st) break;
          if (A.flags & RT_TMJP_F_SUMMARY) {
            st = summary_pass<SP>(A, C, W, lane, site, p_total, scr_seg, X.n_seg, total_len, out7, pt64);
            if (st) break;
            have7 = true;
            if (lane < 7) atomicAdd(&C.sum_acc[lane], out7[lane]);
            if (lane == 7) atomicAdd(&C.sum_acc[7], 1.0);
          }
        }
      }
    } else if (A.mode == RT_TMJP_TRAJ_LOGLIK) {
      const double ll = trajectory_loglik<SP>(A, C, W, lane, p_total, stats_p, pt64);
      if (lane == 0 && A.traj_loglik) A.traj_loglik[traj] = ll;
    } else {   // RT_TMJP_SUMMARY
      st = summary_pass<SP>(A, C, W, lane, site, p_total, scr_seg, X.n_seg, total_len, out7, pt64);
      if (st == 0) {
        have7 = true;
        if (lane < 7) atomicAdd(&C.sum_acc[lane], out7[lane]);
        if (lane == 7) atomicAdd(&C.sum_acc[7], 1.0);
      }
    }
    __syncwarp();
    // ---- shared memory -> trajectory
    if (primary_dirty) {
      for (int v = lane; v < A.n_nodes; v += 32) {
        A.p_node[traj * A.pn_traj_stride + (int64_t)v * A.pn_node_stride] = W.pn[v];
        A.p_cnt[traj * A.pn_traj_stride + (int64_t)v * A.pn_node_stride] = W.pc[v];
      }
      for (int i = lane; i < A.cap_p; i += 32) {
        A.p_time[(size_t)traj * A.cap_p + i] = W.pt[i];
        A.p_sb[(size_t)traj * A.cap_p + i] = W.psb[i];
      }
      if (lane == 0) A.p_total[traj] = p_total;
    }
    if (tol_dirty && A.t_node) {
      for (int v = lane; v < A.n_nodes; v += 32) A.t_node[(size_t)traj * A.n_nodes + v] = W.tn[v];
      for (int i = lane; i < A.n_nodes * NP; i += 32) A.t_cnt[(size_t)traj * A.n_nodes * NP + i] = W.tc[i];
    }
    if (have7 && A.summary_out && lane < 8) A.summary_out[(size_t)traj * 8 + lane] = out7[lane];
    if (st && lane == 0) A.status[traj] = (int8_t)st;
    __syncwarp();
  }
  __syncthreads();
  if (A.prim_dwell && stats_p)
    for (int i = tid; i < S; i += kThreads) if (dwell_s[i] != 0.0) atomicAdd(&A.prim_dwell[i], dwell_s[i]);
  if (A.prim_trans && stats_p)
    for (int i = tid; i < S * S; i += kThreads) if (trans_s[i]) atomicAdd(&A.prim_trans[i], (double)trans_s[i]);
  if (A.tol_stats && stats_t)
    for (int i = tid; i < 4 * NP; i += kThreads) if (tol_s[i] != 0.0) atomicAdd(&A.tol_stats[i], tol_s[i]);
  if (A.summary_sum && tid < 8 && sum_s[tid] != 0.0) atomicAdd(&A.summary_sum[tid], sum_s[tid]);
}

template <int SP, int kWarps>
int launch(const rt_tmjp_args& A, cudaStream_t stream) {
  constexpr int SPAD = SP * 32;
  constexpr int kThreads = kWarps * 32;
  int dev = 0, n_sm = 148;
  RT_CUDA_CHECK(cudaGetDevice(&dev));
  RT_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  Scratch X;
  X.scr_cap = A.cap_p;
  if (A.mode == RT_TMJP_INIT_PRIMARY && A.init_k > 0 && (A.n_nodes - 1) * A.init_k > X.scr_cap)
    X.scr_cap = (A.n_nodes - 1) * A.init_k;
  // tolerance candidates per class and sweep: old toggles + virtual events (<= cap_t), or
  // in init mode one per primary segment
  X.cap_ts = A.cap_t;
  if (A.mode == RT_TMJP_INIT_TOLERANCE) X.cap_ts = (A.n_nodes - 1) + A.cap_p;
  X.n_seg = (A.n_nodes - 1) + A.cap_p;
  const Layout L = make_layout(A.S, SPAD, A.n_parts, A.n_nodes, A.n_ops, A.n_slots, A.cap_p, X.scr_cap, kWarps);
  if (L.total > 220 * 1024) return RT_ERR_UNSUPPORTED;
  auto kern = tmjp_kernel<SP, kWarps>;
  RT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  int per_sm = 1;
  RT_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, L.total));
  if (per_sm < 1) return RT_ERR_UNSUPPORTED;
  int64_t grid = (A.n_traj + kWarps - 1) / kWarps;
  const int64_t max_grid = (int64_t)n_sm * per_sm;
  if (grid > max_grid) grid = max_grid;
  const size_t n_warps = (size_t)grid * kWarps;
  const bool tol = A.n_parts > 0 && A.mode != RT_TMJP_INIT_PRIMARY;
  const bool want_seg = A.n_parts > 0 && (A.mode == RT_TMJP_SUMMARY ||
                                         (A.mode == RT_TMJP_SWEEP && (A.flags & RT_TMJP_F_SUMMARY)));
  const size_t b_beta = al(n_warps * (size_t)X.scr_cap * SPAD * sizeof(double), 256);
  const size_t b_tu = tol ? al(n_warps * (size_t)X.cap_ts * 32 * sizeof(float2), 256) : 0;
  const size_t b_tb = tol ? al(n_warps * (size_t)X.cap_ts * 32 * sizeof(double2), 256) : 0;
  const size_t b_seg = want_seg ? al(n_warps * (size_t)X.n_seg * 32 * sizeof(double2), 256) : 0;
  unsigned char* ws = nullptr;
  RT_CUDA_CHECK(rt_ws_alloc((void**)&ws, b_beta + b_tu + b_tb + b_seg + 256, stream));
  unsigned char* p = ws;
  X.p_beta = reinterpret_cast<double*>(p); p += b_beta;
  X.t_tu = tol ? reinterpret_cast<float2*>(p) : nullptr; p += b_tu;
  X.t_beta = tol ? reinterpret_cast<double2*>(p) : nullptr; p += b_tb;
  X.seg = want_seg ? reinterpret_cast<double2*>(p) : nullptr;
  kern<<<(unsigned)grid, kThreads, L.total, stream>>>(A, X);
  cudaError_t e = cudaGetLastError();
  rt_ws_free(ws, stream);
  if (e != cudaSuccess) { rt_set_last_error(e, __FILE__, __LINE__); return RT_ERR_CUDA; }
  return RT_OK;
}

}  // namespace

template <int SP>
static int launch_fit(const rt_tmjp_args& A, cudaStream_t stream) {
  int rc = launch<SP, 16>(A, stream);
  if (rc == RT_ERR_UNSUPPORTED) rc = launch<SP, 8>(A, stream);
  if (rc == RT_ERR_UNSUPPORTED) rc = launch<SP, 4>(A, stream);
  if (rc == RT_ERR_UNSUPPORTED) rc = launch<SP, 2>(A, stream);
  return rc;
}

int rt_tmjp_dispatch(const rt_tmjp_args& A, cudaStream_t stream) {
  if (A.S <= 32) return launch_fit<1>(A, stream);
  if (A.S <= 64) return launch_fit<2>(A, stream);
  return RT_ERR_UNSUPPORTED;
}
