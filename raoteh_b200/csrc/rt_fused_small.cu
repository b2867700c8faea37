// K2 + K4/K5 FUSED for small state spaces (S <= 4): one persistent kernel prunes a tile of sites
// up the tree and immediately walks the same tile down again -- log-likelihood, posterior
// marginals (on chip) and the per-edge sufficient statistics W_b = sum_sites G_b (x) L_b.
//
// Replaces, batched over sites, the whole per-site chain of
//   _mjp_dense.get_expected_history_statistics   raoteh/sampler/_mjp_dense.py:410-539:
//     pyfelscore.mcy_esd_get_node_to_pmap        (_mcy_dense.py:286-291; spec _mcy.py:611-682)
//     _mc0_dense.get_likelihood                  (_mc0_dense.py:147-212)
//     pyfelscore.mc0_esd_get_node_to_distn       (_mc0_dense.py:381; spec :400-489)
//     pyfelscore.mc0_esd_get_joint_endpoint_distn (_mcy_dense.py:205; spec _mc0_dense.py:217-270)
//     and the joint / conditional accumulation   (_mjp_dense.py:502-510, 521-533).
//
// Why fused: run as two kernels (rt_prune_small.cu with stored partials, then the walk of
// rt_posterior_small.cu) the path writes every internal partial to HBM (n_int * S * 8 B per site,
// 1 KB at BASELINE config C2) and reads it back 0.3 ms later with DRAM latency on a serial
// per-site dependency chain.  Here the partials of a tile go to a per-CTA scratch that the same
// CTA overwrites tile after tile: the lines stay in L2 (the CTAs' scratch together is sized below
// the 126 MB L2), the read-back is an L2 hit, and HBM sees only the observations in and the
// per-site log-likelihood / status out.
//
// The cross-lane reduction of W (S*S values per lane, every edge) goes through a per-warp
// staging tile in shared memory (transposed read, one partial sum per lane, one shuffle) instead
// of the 15-exchange shuffle butterfly of the unfused walk: ~55 instead of ~125 issue slots per
// edge and warp.
#include "rt_common.cuh"
#include "rt_small.cuh"
#include <stdlib.h>

namespace {

using rt_small::ObsVal;
using rt_small::max_hiword;
using rt_small::fast_rcp;

constexpr int kFB = 128;     // threads per CTA
constexpr int kWarps = kFB / 32;
constexpr int kStageRow = 33;   // doubles per staging row: 32 lanes + 1 (conflict-free transposed read)

template <int S> struct PRow { static constexpr int value = (S + 2) & ~1; };   // P row + row sum, even

// the line holding `p` is dead: drop it from L2 without writing it back (p: 128-byte aligned)
// `dep`: a register produced from the last load of that line, so that the discard cannot be issued
// before the load instruction has completed for the whole warp
__device__ __forceinline__ void l2_discard_line(const void* p, double dep) {
  asm volatile("discard.global.L2 [%0], 128;" :: "l"(p), "r"(__double2hiint(dep)) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// Decoded op in shared memory: x = code | flags, y = offset of the edge's P rows in Pt_s (doubles),
// z = offset of the edge's block in W_s, w = offset of the node's partial in the CTA's scratch (doubles; MSG_SLOT: the child,
// STORE: the node itself).
template <int S, int OBS, bool PACKED, int NS>
__global__ void __launch_bounds__(kFB, 4)
fused_small_kernel(int64_t n_sites, int64_t stride, const int4* __restrict__ program, int n_ops,
                   int n_slots, int n_nodes, int n_store, const double* __restrict__ P,
                   const double* __restrict__ root_distn, const void* __restrict__ obs,
                   double* __restrict__ scratch, double* __restrict__ loglik,
                   int8_t* __restrict__ status, double* __restrict__ loglik_sum,
                   double* __restrict__ W, double* __restrict__ root_post_sum, int use_discard) {
  constexpr int obs_packed = PACKED ? 1 : 0;
  constexpr int SS = S * S;
  constexpr int PR = PRow<S>::value;
  constexpr int V = 16;                 // staged values per lane (S*S <= 16)
  constexpr int TS = NS * kFB;          // sites per tile
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int4* prog_s = reinterpret_cast<int4*>(smem_raw);
  double* Pt_s = reinterpret_cast<double*>(prog_s + n_ops);   // [n_nodes][S][PR]: row of P, then its sum
  double* W_s = Pt_s + (size_t)n_nodes * S * PR;        // [n_nodes][SS]
  double* stage_s = W_s + (size_t)n_nodes * SS;         // [kWarps][V][kStageRow]
  double* fresh_s = stage_s + kWarps * V * kStageRow;   // [NS][S][kFB]: the partial stored by the previous op
  double* pi_s = fresh_s + NS * S * kFB;                // [4]
  double* zero_s = pi_s + 4;            // [2]
  double* rp_s = zero_s + 2;            // [4]
  long long* oo_s = reinterpret_cast<long long*>(rp_s + 4);         // [n_ops] own obs-row offset
  long long* pre_s = oo_s + n_ops;                                  // [n_ops + 1] next obs row after ip

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n_ops; i += kFB) {
    const int4 op = program[i];
    const int code = op.x & 0xff;
    int4 d;
    d.x = op.x;
    d.y = op.y * S * PR;
    d.z = op.y * SS;
    d.w = op.w * S * TS;
    prog_s[i] = d;
    oo_s[i] = (code == OP_MSG_OBS || code == OP_APPLY_OBS) ? ObsVal<S, OBS>::row_off(op.z, stride, obs_packed) : -1;
  }
  __syncthreads();
  if (tid == 0) {
    long long nxt = -1;
    for (int i = n_ops - 1; i >= 0; --i) {
      pre_s[i] = nxt;
      if (oo_s[i] >= 0) nxt = oo_s[i];
    }
    pre_s[n_ops] = nxt;
  }
  if (tid < S) { pi_s[tid] = root_distn ? root_distn[tid] : 1.0; rp_s[tid] = 0.0; }
  if (tid < 2) zero_s[tid] = 0.0;
  for (int i = tid; i < n_nodes * S; i += kFB) {
    double t = 0.0;
#pragma unroll
    for (int b = 0; b < S; ++b) { const double v = P[(size_t)i * S + b]; Pt_s[i * PR + b] = v; t += v; }
    Pt_s[i * PR + S] = t;
    if (PR > S + 1) Pt_s[i * PR + S + 1] = 0.0;
  }
  for (int i = tid; i < n_nodes * SS; i += kFB) W_s[i] = 0.0;
  __syncthreads();

  double* scr = scratch + (size_t)blockIdx.x * n_store * S * TS + tid;    // + op.w + a*TS + q*kFB
  double* stg = stage_s + warp * V * kStageRow;
  double* fr = fresh_s + tid;                                             // + (q*S + a)*kFB
  const bool discard_lane = use_discard && (lane & 15) == 0;              // one lane per 128-byte line
  const int64_t tiles = (n_sites + TS - 1) / TS;
  double my_ll = 0.0;

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t site0 = tile * TS + tid;
    int64_t site[NS];
    bool active[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      site[q] = site0 + q * kFB;
      active[q] = site[q] < n_sites;
      if (!active[q]) site[q] = n_sites - 1;      // clamp: loads stay in range, results discarded
    }
    // the NEXT tile's code rows: pulled into L2 now, a whole tile time ahead of their use
    if (OBS == OBS_CODES && tile + gridDim.x < tiles) {
      constexpr int row_bytes = PACKED ? TS / 2 : TS;
      constexpr int LPR = (row_bytes + 127) / 128;
      const int64_t nb = (tile + gridDim.x) * (int64_t)row_bytes;
      for (int i = tid; i < n_ops * LPR; i += kFB) {
        const long long off = oo_s[i / LPR];
        const int64_t a = nb + (i % LPR) * 128;
        if (off >= 0 && a < (PACKED ? (n_sites + 1) / 2 : n_sites))
          l2_prefetch(reinterpret_cast<const uint8_t*>(obs) + off + a);
      }
    }
    double cur[NS][S];
    // ================================== UP ==================================
    {
      double acc[NS][S];
      int etot[NS];
      ObsVal<S, OBS> ov[NS];
      const void* obs_q[NS];
      int nib[NS];
#pragma unroll
      for (int q = 0; q < NS; ++q) {
#pragma unroll
        for (int a = 0; a < S; ++a) acc[q][a] = 1.0;
        etot[q] = 0;
        obs_q[q] = ObsVal<S, OBS>::base(obs, site[q], obs_packed);
        nib[q] = (int)(site[q] & 1) * 4;
        ov[q].init();
        ov[q].fetch(obs_q[q], pre_s[n_ops], stride);
      }
      int4 op_n = prog_s[0];
      for (int ip = 0; ip < n_ops; ++ip) {
        const int4 op = op_n;
        if (ip + 1 < n_ops) op_n = prog_s[ip + 1];      // next op's word arrives under this op's work
        const int code = op.x & 0xff;
        const double* Pc = Pt_s + op.y;
        auto matvec = [&](const double (&v)[NS][S], double (&m)[NS][S]) {
#pragma unroll
          for (int a = 0; a < S; ++a) {
#pragma unroll
            for (int q = 0; q < NS; ++q) m[q][a] = 0.0;
            if constexpr (S % 2 == 0) {
              const double2* pr = reinterpret_cast<const double2*>(Pc + a * PR);
#pragma unroll
              for (int b = 0; b < S / 2; ++b) {
                const double2 p = pr[b];
#pragma unroll
                for (int q = 0; q < NS; ++q) {
                  m[q][a] = fma(p.x, v[q][2 * b], m[q][a]);
                  m[q][a] = fma(p.y, v[q][2 * b + 1], m[q][a]);
                }
              }
            } else {
#pragma unroll
              for (int b = 0; b < S; ++b) {
                const double p = Pc[a * PR + b];
#pragma unroll
                for (int q = 0; q < NS; ++q) m[q][a] = fma(p, v[q][b], m[q][a]);
              }
            }
          }
        };
        switch (code) {
          case OP_MSG_SLOT: {
            double v[NS][S], m[NS][S];
            if (op.x & OP_FLAG_FRESH) {        // stored by the previous op: still in shared memory
#pragma unroll
              for (int q = 0; q < NS; ++q)
#pragma unroll
                for (int b = 0; b < S; ++b) v[q][b] = fr[(q * S + b) * kFB];
            } else {                           // an older sibling: back from the L2 scratch
              const double* src = scr + op.w;
#pragma unroll
              for (int q = 0; q < NS; ++q)
#pragma unroll
                for (int b = 0; b < S; ++b) v[q][b] = __ldcg(src + b * TS + q * kFB);
            }
            matvec(v, m);
#pragma unroll
            for (int q = 0; q < NS; ++q)
#pragma unroll
              for (int a = 0; a < S; ++a) acc[q][a] *= m[q][a];
          } break;
          case OP_MSG_OBS: {
            if constexpr (OBS == OBS_CODES) {
#pragma unroll
              for (int q = 0; q < NS; ++q) {
                const int k = ov[q].code(obs_packed ? nib[q] : -1);
                // branch-free: column k of P, the row sums (unobserved), or zeros (invalid code)
                const bool miss = k == RT_MISSING;
                const double* src = miss ? (Pc + S) : (k < S ? Pc + k : zero_s);
                const int strd = (miss || k < S) ? PR : 0;
#pragma unroll
                for (int a = 0; a < S; ++a) acc[q][a] *= src[a * strd];
              }
            } else {
              double v[NS][S], m[NS][S];
#pragma unroll
              for (int q = 0; q < NS; ++q)
#pragma unroll
                for (int b = 0; b < S; ++b) v[q][b] = ov[q].get(b, -1);
              matvec(v, m);
#pragma unroll
              for (int q = 0; q < NS; ++q)
#pragma unroll
                for (int a = 0; a < S; ++a) acc[q][a] *= m[q][a];
            }
            const long long off = pre_s[ip];
#pragma unroll
            for (int q = 0; q < NS; ++q) ov[q].fetch(obs_q[q], off, stride);
          } break;
          case OP_MSG_ONES: {
#pragma unroll
            for (int a = 0; a < S; ++a) {
              const double r = Pc[a * PR + S];
#pragma unroll
              for (int q = 0; q < NS; ++q) acc[q][a] *= r;
            }
          } break;
          case OP_APPLY_OBS: {
#pragma unroll
            for (int q = 0; q < NS; ++q)
#pragma unroll
              for (int a = 0; a < S; ++a) acc[q][a] *= ov[q].get(a, obs_packed ? nib[q] : -1);
            const long long off = pre_s[ip];
#pragma unroll
            for (int q = 0; q < NS; ++q) ov[q].fetch(obs_q[q], off, stride);
          } break;
          case OP_STORE:
          case OP_ROOT: {
#pragma unroll
            for (int q = 0; q < NS; ++q) {
              const int hmx = max_hiword<S>(acc[q]);
              if (hmx >= 0x00100000) {        // largest entry is a positive normal number
                const int e = (hmx >> 20) - 1023;
                const double sc = rt_pow2_neg(e);
#pragma unroll
                for (int a = 0; a < S; ++a) acc[q][a] *= sc;
                etot[q] += e;
              }
              if (code == OP_STORE) {
                double* pp = scr + op.w + q * kFB;
                const bool keep = (op.x & OP_FLAG_KEEP_ON_CHIP) != 0;
#pragma unroll
                for (int a = 0; a < S; ++a) {
                  __stcg(pp + a * TS, acc[q][a]);      // L2-resident per-CTA scratch
                  if (keep) fr[(q * S + a) * kFB] = acc[q][a];
                  acc[q][a] = 1.0;
                }
              } else {
                // root combine (_mc0_dense.py:147-212) and the root marginal (_mc0_dense.py:400-489)
                double lk = 0.0;
#pragma unroll
                for (int a = 0; a < S; ++a) { cur[q][a] = pi_s[a] * acc[q][a]; lk += cur[q][a]; }
                const bool ok = active[q] && lk > 0.0;
                if (active[q]) {
                  if (lk > 0.0) {
                    const double ll = log(lk) + (double)etot[q] * RT_LN2;
                    loglik[site[q]] = ll;
                    status[site[q]] = RT_SITE_OK;
                    my_ll += ll;
                  } else {
                    loglik[site[q]] = -INFINITY;
                    status[site[q]] = RT_SITE_STRUCTURAL_ZERO;
                  }
                }
                // a dead lane (beyond the last site, or zero likelihood) walks down with a zero
                // marginal: every G, D and W contribution it computes is exactly zero
                const double inv = ok ? 1.0 / lk : 0.0;
#pragma unroll
                for (int a = 0; a < S; ++a) cur[q][a] *= inv;
              }
            }
          } break;
          default: break;
        }
      }
    }
    // posterior of the root, summed over sites
    if (root_post_sum) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < NS; ++q) t += cur[q][s];
        t = rt_warp_sum(t);
        if (lane == 0 && t != 0.0) atomicAdd(&rp_s[s], t);
      }
    }

    // ================================= DOWN =================================
    // The walk of one site is a serial chain, so the loads of op ip-1 (stored partial of an internal
    // child from the L2 scratch, or the code byte of a leaf) are issued into a second register
    // buffer before op ip is computed: ping-pong buffers, loop unrolled by two.
    auto issue = [&](int j, double (&Lb)[NS][S], int (&kb)[NS]) {
      if (j < 0) return;
      const int4 nx = prog_s[j];
      const int ncode = nx.x & 0xff;
      if (ncode == OP_MSG_SLOT) {
        const double* src = scr + nx.w;
#pragma unroll
        for (int q = 0; q < NS; ++q)
#pragma unroll
          for (int s = 0; s < S; ++s) Lb[q][s] = __ldcg(src + s * TS + q * kFB);
      } else if (ncode == OP_MSG_OBS && OBS == OBS_CODES) {
        const uint8_t* row = reinterpret_cast<const uint8_t*>(obs) + oo_s[j];
#pragma unroll
        for (int q = 0; q < NS; ++q) {
          int k;
          if (obs_packed) {
            k = (row[site[q] >> 1] >> ((int)(site[q] & 1) * 4)) & 15;
            if (k == 15) k = RT_MISSING;
          } else {
            k = row[site[q]];
          }
          kb[q] = k;
        }
      }
    };
    auto do_op = [&](int ip, const double (&Lb)[NS][S], const int (&kb)[NS]) {
      const int4 op = prog_s[ip];
      const int code = op.x & 0xff;
      if (code == OP_STORE) {
        // a marginal parked in the scratch by a non-fresh OP_MSG_SLOT (a fresh one is still in `cur`)
        if (op.x & OP_FLAG_PARK) {
          const double* src = scr + op.w;
#pragma unroll
          for (int q = 0; q < NS; ++q)
#pragma unroll
            for (int s = 0; s < S; ++s) {
              cur[q][s] = __ldcg(src + s * TS + q * kFB);
              if (discard_lane) l2_discard_line(src + s * TS + q * kFB, cur[q][s]);
            }
        }
        return;
      }
      if (code > OP_MSG_ONES) return;       // OP_APPLY_OBS / OP_ROOT: nothing to do on the way down
      const double* Pc = Pt_s + op.y;
      double w[V];
#pragma unroll
      for (int i = 0; i < V; ++i) w[i] = 0.0;
      if (code == OP_MSG_SLOT) {
        // ---- internal child: G = D_parent / (P L), D_child = L o (P^T G), W += G (x) L ----
        double Pr[S][S];
#pragma unroll
        for (int a = 0; a < S; ++a) {
          if constexpr (S % 2 == 0) {
#pragma unroll
            for (int b = 0; b < S / 2; ++b) {
              const double2 p = reinterpret_cast<const double2*>(Pc + a * PR)[b];
              Pr[a][2 * b] = p.x; Pr[a][2 * b + 1] = p.y;
            }
          } else {
#pragma unroll
            for (int b = 0; b < S; ++b) Pr[a][b] = Pc[a * PR + b];
          }
        }
        const bool fresh = (op.x & OP_FLAG_FRESH) != 0;
        double* dst = scr + op.w;
#pragma unroll
        for (int q = 0; q < NS; ++q) {
          double G[S];
#pragma unroll
          for (int a = 0; a < S; ++a) {
            double m = 0.0;
#pragma unroll
            for (int b = 0; b < S; ++b) m = fma(Pr[a][b], Lb[q][b], m);
            G[a] = (m > 0.0) ? cur[q][a] * fast_rcp(m) : 0.0;
          }
#pragma unroll
          for (int i = 0; i < SS; ++i) w[i] = fma(G[i / S], Lb[q][i % S], w[i]);
#pragma unroll
          for (int b = 0; b < S; ++b) {
            double t = 0.0;
#pragma unroll
            for (int a = 0; a < S; ++a) t = fma(G[a], Pr[a][b], t);
            t *= Lb[q][b];
            // the child's marginal: kept in registers when the child's own ops come next, else
            // parked in the scratch IN PLACE of its partial (which is dead from here on)
            if (fresh) {
              cur[q][b] = t;
              if (discard_lane) l2_discard_line(dst + b * TS + q * kFB, t);
            } else {
              __stcg(dst + b * TS + q * kFB, t);
            }
          }
        }
      } else if (code == OP_MSG_OBS && OBS == OBS_CODES) {
        // ---- observed leaf, hard code k: P L is column k of P (the row sums when unobserved) ----
#pragma unroll
        for (int q = 0; q < NS; ++q) {
          const int k = kb[q];
          const bool miss = k == RT_MISSING;
          const int col = (miss || k >= S) ? S : k;
          bool hit[S];
#pragma unroll
          for (int b = 0; b < S; ++b) hit[b] = miss || k == b;
#pragma unroll
          for (int a = 0; a < S; ++a) {
            const double m = Pc[a * PR + col];
            const double g = (m > 0.0) ? cur[q][a] * fast_rcp(m) : 0.0;
#pragma unroll
            for (int b = 0; b < S; ++b)
              if (hit[b]) w[a * S + b] += g;
          }
        }
      } else {
        // ---- leaf with a mask / dense emission row / no observation ----
#pragma unroll
        for (int q = 0; q < NS; ++q) {
          double L[S];
          if (code == OP_MSG_ONES) {
#pragma unroll
            for (int s = 0; s < S; ++s) L[s] = 1.0;
          } else if (OBS == OBS_MASK) {
            const unsigned long long mk =
                reinterpret_cast<const unsigned long long*>(obs)[oo_s[ip] + site[q]];
#pragma unroll
            for (int s = 0; s < S; ++s) L[s] = ((mk >> s) & 1ull) ? 1.0 : 0.0;
          } else {
            const double* d = reinterpret_cast<const double*>(obs) + oo_s[ip] + site[q];
#pragma unroll
            for (int s = 0; s < S; ++s) L[s] = d[(int64_t)s * stride];
          }
#pragma unroll
          for (int a = 0; a < S; ++a) {
            double m = 0.0;
#pragma unroll
            for (int b = 0; b < S; ++b) m = fma(Pc[a * PR + b], L[b], m);
            const double g = (m > 0.0) ? cur[q][a] * fast_rcp(m) : 0.0;
#pragma unroll
            for (int b = 0; b < S; ++b) w[a * S + b] = fma(g, L[b], w[a * S + b]);
          }
        }
      }
      // ---- W_c += sum over the warp's 32*NS sites: transpose through the warp's staging tile ----
      __syncwarp();                          // the previous edge's readers are done
#pragma unroll
      for (int i = 0; i < V; ++i) stg[i * kStageRow + lane] = w[i];
      __syncwarp();
      {
        // lane (o, h) sums half h of row o; row stride 33 doubles puts the 32 lanes of one load
        // on 16 distinct bank pairs with compile-time offsets
        const int o = lane >> 1, h = lane & 1;
        const double* row = stg + o * kStageRow + h * 16;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          s0 += row[j]; s1 += row[j + 1]; s2 += row[j + 2]; s3 += row[j + 3];
        }
        double sum = (s0 + s1) + (s2 + s3);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        if (h == 0 && o < SS && sum != 0.0) atomicAdd(&W_s[op.z + o], sum);
      }
    };
    double LA[NS][S], LB[NS][S];
    int kA[NS], kB[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      kA[q] = 0; kB[q] = 0;
#pragma unroll
      for (int s = 0; s < S; ++s) { LA[q][s] = 0.0; LB[q][s] = 0.0; }
    }
    int ip = n_ops - 2;                      // n_ops - 1 is OP_ROOT, done above
    issue(ip, LA, kA);
    for (; ip >= 1; ip -= 2) {
      issue(ip - 1, LB, kB);
      do_op(ip, LA, kA);
      issue(ip - 2, LA, kA);
      do_op(ip - 1, LB, kB);
    }
    if (ip == 0) do_op(0, LA, kA);
  }

  __syncthreads();
  for (int i = tid; i < n_nodes * SS; i += kFB) {
    const double v = W_s[i];
    // restricted to P > 0 (_mjp_dense.py:505-510 divides the joint by P where it is positive)
    if (v != 0.0 && Pt_s[(i / SS) * S * PR + ((i % SS) / S) * PR + (i % S)] > 0.0) atomicAdd(&W[i], v);
  }
  if (root_post_sum && tid < S && rp_s[tid] != 0.0) atomicAdd(&root_post_sum[tid], rp_s[tid]);
  if (loglik_sum) {
    __shared__ double red[kWarps];
    const double ws = rt_warp_sum(my_ll);
    if (lane == 0) red[warp] = ws;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int i = 0; i < kWarps; ++i) t += red[i];
      if (t != 0.0) atomicAdd(loglik_sum, t);
    }
  }
}

template <int S, int OBS, bool PACKED>
int launch(int64_t n_sites, int64_t stride, const int4* program, int n_ops, int n_slots,
           int n_nodes, int n_store, const double* P, const double* root_distn, const void* obs,
           double* loglik, int8_t* status, double* loglik_sum, double* W, double* root_post_sum,
           int ctas_per_sm, cudaStream_t stream, bool* handled) {
  constexpr int NS = 2;
  // RT_FUSED_DISCARD=1: drop dead scratch lines from L2 with discard.global.L2 right after their
  // last read (measured slower on B200: 1.155 ms against 1.00 ms at C2 size; default off)
  static const int use_discard = [] { const char* e = getenv("RT_FUSED_DISCARD"); return (e && e[0] == '1') ? 1 : 0; }();
  constexpr int PR = PRow<S>::value;
  auto kern = fused_small_kernel<S, OBS, PACKED, NS>;
  const size_t smem = (sizeof(int4) + 2 * sizeof(long long)) * (size_t)n_ops + 32 +
                      sizeof(double) * (4 + 2 + 4) +
                      sizeof(double) * (size_t)n_nodes * (S * PR + S * S) +
                      sizeof(double) * kWarps * 16 * kStageRow +
                      sizeof(double) * (size_t)NS * S * kFB;
  *handled = false;
  if (smem > 110 * 1024) return RT_OK;         // caller falls back to the two-kernel path
  RT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  RT_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFB, smem));
  if (per_sm < 1) per_sm = 1;
  if (ctas_per_sm > 0 && ctas_per_sm < per_sm) per_sm = ctas_per_sm;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (n_sites + kFB * NS - 1) / (kFB * NS);
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > tiles) grid = tiles;
  double* scratch = nullptr;
  const size_t scr_bytes = sizeof(double) * (size_t)grid * n_store * S * NS * kFB;
  RT_CUDA_CHECK(rt_ws_alloc((void**)&scratch, scr_bytes, stream));
  kern<<<(unsigned)grid, kFB, smem, stream>>>(n_sites, stride, program, n_ops, n_slots, n_nodes,
                                              n_store, P, root_distn, obs, scratch, loglik, status,
                                              loglik_sum, W, root_post_sum, use_discard);
  cudaError_t e = cudaGetLastError();
  rt_ws_free(scratch, stream);
  RT_CUDA_CHECK(e);
  *handled = true;
  return RT_OK;
}

template <int S>
int run(int obs_kind, int64_t n_sites, int64_t stride, const int4* program, int n_ops, int n_slots,
        int n_nodes, int n_store, const double* P, const double* root_distn, const void* obs,
        double* loglik, int8_t* status, double* loglik_sum, double* W, double* root_post_sum,
        int ctas_per_sm, cudaStream_t stream, bool* handled) {
#define RT_ARGS n_sites, stride, program, n_ops, n_slots, n_nodes, n_store, P, root_distn, obs, loglik, \
                status, loglik_sum, W, root_post_sum, ctas_per_sm, stream, handled
  switch (obs_kind) {
    case OBS_CODES: return launch<S, OBS_CODES, false>(RT_ARGS);
    case 3: return launch<S, OBS_CODES, true>(RT_ARGS);
    case OBS_MASK: return launch<S, OBS_MASK, false>(RT_ARGS);
    case OBS_DENSE: return launch<S, OBS_DENSE, false>(RT_ARGS);
  }
#undef RT_ARGS
  return RT_ERR_ARG;
}

}  // namespace

// handled = false (and RT_OK): shape not covered (S > 4 or shared-memory budget) -- the caller runs
// the two-kernel path instead.
int rt_fused_small_dispatch(int S, int obs_kind, int64_t n_sites, int64_t stride,
                            const int32_t* program, int n_ops, int n_slots, int n_nodes, int n_store,
                            const double* P, const double* root_distn, const void* obs,
                            double* loglik, int8_t* status, double* loglik_sum, double* W,
                            double* root_post_sum, int ctas_per_sm, cudaStream_t stream, bool* handled) {
  const int4* prog = reinterpret_cast<const int4*>(program);
#define RT_ARGS obs_kind, n_sites, stride, prog, n_ops, n_slots, n_nodes, n_store, P, root_distn, obs, \
                loglik, status, loglik_sum, W, root_post_sum, ctas_per_sm, stream, handled
  *handled = false;
  switch (S) {
    case 2: return run<2>(RT_ARGS);
    case 3: return run<3>(RT_ARGS);
    case 4: return run<4>(RT_ARGS);
  }
#undef RT_ARGS
  return RT_OK;
}
