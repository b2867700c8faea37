// K2: Felsenstein pruning for small state spaces (S <= 8), one thread per site.
//
// Replaces, batched over sites, the reference's per-site chain
//   pyfelscore.mcy_esd_get_node_to_pmap  (called raoteh/sampler/_mcy_dense.py:286-291,
//   spec _mcy.py:611-682, emissions _mcz.py:138-163)
//   + _mc0_dense.get_likelihood (raoteh/sampler/_mc0_dense.py:147-212).
//
// Each thread walks the host-built upward program (lowering.TreeSchedule.
// up_program) for its site.  Partials of nodes whose parent has not been
// processed yet live in a per-thread stack in shared memory
// ([slot][state][thread], conflict-free); transition matrices are staged once
// per CTA in shared memory and read as warp-uniform broadcasts.  Every stored
// partial is rescaled by an exact power of two and the integer exponent is
// carried, so log-likelihoods never underflow and the scaling adds no rounding.
// HBM traffic per site: the observation row(s) in, 8 B log-lik + 1 B status out
// (+ the stored partials when the posterior pass will follow).
#include "rt_common.cuh"

namespace {

constexpr int kBlock = 128;

// One observation row of one site, fetched one obs-consuming op ahead of its use
// (one specialisation per encoding so that only the live member occupies registers).
template <int S, int OBS> struct ObsVal;

template <int S> struct ObsVal<S, OBS_CODES> {
  int k;
  __device__ __forceinline__ void init() { k = RT_MISSING; }
  // packed: two codes per byte (RT_OBS_CODES4), site i in nibble i & 1 of byte i >> 1, 15 = unobserved
  __device__ __forceinline__ void fetch(const void* __restrict__ obs, int row, int64_t stride, int64_t site,
                                        int packed) {
    if (row < 0) return;
    if (packed) {
      const int b = reinterpret_cast<const uint8_t*>(obs)[(int64_t)row * ((stride + 1) >> 1) + (site >> 1)];
      k = (b >> ((int)(site & 1) * 4)) & 15;
      if (k == 15) k = RT_MISSING;
    } else {
      k = reinterpret_cast<const uint8_t*>(obs)[(int64_t)row * stride + site];
    }
  }
  __device__ __forceinline__ double get(int b) const { return (k == RT_MISSING || k == b) ? 1.0 : 0.0; }
};

template <int S> struct ObsVal<S, OBS_MASK> {
  unsigned long long mk;
  __device__ __forceinline__ void init() { mk = ~0ull; }
  __device__ __forceinline__ void fetch(const void* __restrict__ obs, int row, int64_t stride, int64_t site,
                                        int) {
    if (row >= 0) mk = reinterpret_cast<const unsigned long long*>(obs)[(int64_t)row * stride + site];
  }
  __device__ __forceinline__ double get(int b) const { return ((mk >> b) & 1ull) ? 1.0 : 0.0; }
};

template <int S> struct ObsVal<S, OBS_DENSE> {
  double d[S];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < S; ++s) d[s] = 1.0;
  }
  __device__ __forceinline__ void fetch(const void* __restrict__ obs, int row, int64_t stride, int64_t site,
                                        int) {
    if (row < 0) return;
    const double* p = reinterpret_cast<const double*>(obs) + (int64_t)row * S * stride + site;
#pragma unroll
    for (int s = 0; s < S; ++s) d[s] = __ldcs(p + (int64_t)s * stride);
  }
  __device__ __forceinline__ double get(int b) const { return d[b]; }
};

// Optional: dense emission rows streamed through a per-thread ring in shared memory with
// cp.async (RT_DENSE_RING rows in flight per site).  MEASURED SLOWER on B200 than the
// one-row register prefetch (C2, 1e6 sites: ring 0/1/2/4 -> 0.287/0.360/0.392/0.524 ms): the
// ring's shared memory costs more occupancy than the extra bytes in flight buy, and the
// kernel's floor is its instruction issue (0.262 ms with 1-byte codes), so it is off (0).
#ifndef RT_DENSE_RING
#define RT_DENSE_RING 0
#endif
constexpr int kRing = RT_DENSE_RING;            // rows in flight
constexpr int kRingSlots = kRing + 1;

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

// decoded op in shared memory: x = opcode, y = P offset in doubles (node*S*S),
// z = stack offset in doubles (slot*NS*S*kBlock) or obs row, w = store index;
// pre_s[ip] = observation row to prefetch when op ip is reached (-1 none).
// NS sites per thread: the op decode and every P element are loaded once and used
// for NS sites (the kernel is issue-bound, not FLOP-bound), and NS independent
// dependency chains hide the FP64 latency.
template <int S, int OBS, bool STORE, bool P_SMEM, int NS>
__global__ void __launch_bounds__(kBlock)
prune_small_kernel(int64_t n_sites, int64_t stride,
                   const int4* __restrict__ program, int n_ops, int n_slots, int n_nodes,
                   const double* __restrict__ P,
                   const double* __restrict__ root_distn,
                   const void* __restrict__ obs,
                   double* __restrict__ partials, int32_t* __restrict__ exponents,
                   double* __restrict__ loglik, int8_t* __restrict__ status,
                   double* __restrict__ loglik_sum, int obs_packed) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // carve: decoded program | prefetch rows | pi | rowsum | P (optional) | stack | stack exponents
  int4* prog_s = reinterpret_cast<int4*>(smem_raw);
  int* pre_s = reinterpret_cast<int*>(prog_s + n_ops);
  double* pi_s = reinterpret_cast<double*>(pre_s + ((n_ops + 4) & ~3));
  double* rowsum_s = pi_s + S;
  double* P_s = rowsum_s + n_nodes * S;
  double* stk = P_s + (P_SMEM ? n_nodes * S * S : 0);
  int* estk = reinterpret_cast<int*>(stk + n_slots * NS * S * kBlock);
  // dense emissions only: ring [slot][site q][state][thread] and the obs rows in order of use
  double* ring = reinterpret_cast<double*>(
      (reinterpret_cast<uintptr_t>(estk + n_slots * NS * kBlock) + 7) & ~(uintptr_t)7);
  int* orow_s = reinterpret_cast<int*>(ring + ((OBS == OBS_DENSE && kRing > 0) ? kRingSlots * NS * S * kBlock : 0));

  const int tid = threadIdx.x;
  for (int i = tid; i < n_ops; i += kBlock) {
    int4 op = program[i];
    const int code = op.x & 0xff;
    int4 d;
    d.x = code;
    d.y = op.y * S * S;
    d.z = (code == OP_MSG_SLOT || code == OP_STORE) ? op.z * NS * S * kBlock : op.z;
    d.w = op.w;
    prog_s[i] = d;
  }
  __syncthreads();
  if (tid == 0) {
    // pre_s[ip]: row consumed by the next obs-consuming op after ip
    int nxt = -1;
    for (int i = n_ops - 1; i >= 0; --i) {
      pre_s[i] = nxt;
      const int code = prog_s[i].x;
      if (code == OP_MSG_OBS || code == OP_APPLY_OBS) nxt = prog_s[i].z;
    }
    pre_s[n_ops] = nxt;     // first row of the program
    if (OBS == OBS_DENSE && kRing > 0) {
      int j = 0;
      for (int i = 0; i < n_ops; ++i) {
        const int code = prog_s[i].x;
        if (code == OP_MSG_OBS || code == OP_APPLY_OBS) orow_s[j++] = prog_s[i].z;
      }
      for (int d = 0; d < kRing; ++d) orow_s[j++] = -1;
    }
  }
  if (tid < S) pi_s[tid] = root_distn ? root_distn[tid] : 1.0;
  for (int i = tid; i < n_nodes * S; i += kBlock) {
    double r = 0.0;
#pragma unroll
    for (int b = 0; b < S; ++b) r += P[(size_t)i * S + b];
    rowsum_s[i] = r;
  }
  if (P_SMEM)
    for (int i = tid; i < n_nodes * S * S; i += kBlock) P_s[i] = P[i];
  __syncthreads();

  int64_t site[NS];
  bool active[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    site[q] = ((int64_t)blockIdx.x * NS + q) * kBlock + tid;
    active[q] = site[q] < n_sites;
    if (!active[q]) site[q] = n_sites - 1;      // clamp: loads stay in range, results discarded
  }
  double my_ll = 0.0;

  {
    double acc[NS][S];
    int esum[NS];
    ObsVal<S, OBS> cur[NS], nxt[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
#pragma unroll
      for (int a = 0; a < S; ++a) acc[q][a] = 1.0;
      esum[q] = 0;
      nxt[q].init();
      cur[q].init();
      if (!(OBS == OBS_DENSE && kRing > 0)) nxt[q].fetch(obs, pre_s[n_ops], stride, site[q], obs_packed);
    }
    int jobs = 0;                    // index of the next obs-consuming op (dense ring)
    auto ring_issue = [&](int j) {   // start the copy of the j-th obs row into its ring slot
      const int row = orow_s[j];
      if (row >= 0) {
        double* dst = ring + (size_t)(j % kRingSlots) * NS * S * kBlock + tid;
#pragma unroll
        for (int q = 0; q < NS; ++q) {
          const double* src = reinterpret_cast<const double*>(obs) + (int64_t)row * S * stride + site[q];
#pragma unroll
          for (int b = 0; b < S; ++b) cp_async8(dst + (q * S + b) * kBlock, src + (int64_t)b * stride);
        }
      }
      cp_async_commit();
    };
    if (OBS == OBS_DENSE && kRing > 0) {
#pragma unroll
      for (int d = 0; d < kRing; ++d) ring_issue(d);
    }

    for (int ip = 0; ip < n_ops; ++ip) {
      const int4 op = prog_s[ip];
      const double* Pc = P_SMEM ? (P_s + op.y) : (P + op.y);
      if (op.x == OP_MSG_OBS || op.x == OP_APPLY_OBS) {
        if constexpr (OBS == OBS_DENSE && kRing > 0) {
          cp_async_wait<(kRing > 0 ? kRing - 1 : 0)>();          // the oldest row in flight has landed
          const double* src = ring + (size_t)(jobs % kRingSlots) * NS * S * kBlock + tid;
#pragma unroll
          for (int q = 0; q < NS; ++q)
#pragma unroll
            for (int b = 0; b < S; ++b) cur[q].d[b] = src[(q * S + b) * kBlock];
          ring_issue(jobs + kRing);            // reuses the slot read one obs op ago
          ++jobs;
        } else {
          const int row = pre_s[ip];
#pragma unroll
          for (int q = 0; q < NS; ++q) {
            cur[q] = nxt[q];
            nxt[q].fetch(obs, row, stride, site[q], obs_packed);
          }
        }
      }
      switch (op.x) {
        case OP_MSG_SLOT: {
          double v[NS][S], m[NS][S];
#pragma unroll
          for (int q = 0; q < NS; ++q) {
            const double* sp = stk + op.z + q * S * kBlock + tid;
#pragma unroll
            for (int b = 0; b < S; ++b) v[q][b] = sp[b * kBlock];
            esum[q] += estk[op.z / S + q * kBlock + tid];
#pragma unroll
            for (int a = 0; a < S; ++a) m[q][a] = 0.0;
          }
#pragma unroll
          for (int a = 0; a < S; ++a)
#pragma unroll
            for (int b = 0; b < S; ++b) {
              const double p = Pc[a * S + b];
#pragma unroll
              for (int q = 0; q < NS; ++q) m[q][a] = fma(p, v[q][b], m[q][a]);
            }
#pragma unroll
          for (int q = 0; q < NS; ++q)
#pragma unroll
            for (int a = 0; a < S; ++a) acc[q][a] *= m[q][a];
        } break;
        case OP_MSG_OBS: {
          if constexpr (OBS == OBS_CODES) {
#pragma unroll
            for (int q = 0; q < NS; ++q) {
              const int k = cur[q].k;
              if (k == RT_MISSING) {
#pragma unroll
                for (int a = 0; a < S; ++a) acc[q][a] *= rowsum_s[op.y / S + a];
              } else if (k < S) {
#pragma unroll
                for (int a = 0; a < S; ++a) acc[q][a] *= Pc[a * S + k];
              } else {
#pragma unroll
                for (int a = 0; a < S; ++a) acc[q][a] = 0.0;
              }
            }
          } else {
            double v[NS][S], m[NS][S];
#pragma unroll
            for (int q = 0; q < NS; ++q)
#pragma unroll
              for (int b = 0; b < S; ++b) {
                v[q][b] = cur[q].get(b);
                m[q][b] = 0.0;
              }
#pragma unroll
            for (int a = 0; a < S; ++a)
#pragma unroll
              for (int b = 0; b < S; ++b) {
                const double p = Pc[a * S + b];
#pragma unroll
                for (int q = 0; q < NS; ++q) m[q][a] = fma(p, v[q][b], m[q][a]);
              }
#pragma unroll
            for (int q = 0; q < NS; ++q)
#pragma unroll
              for (int a = 0; a < S; ++a) acc[q][a] *= m[q][a];
          }
        } break;
        case OP_MSG_ONES: {
#pragma unroll
          for (int a = 0; a < S; ++a) {
            const double r = rowsum_s[op.y / S + a];
#pragma unroll
            for (int q = 0; q < NS; ++q) acc[q][a] *= r;
          }
        } break;
        case OP_APPLY_OBS: {
#pragma unroll
          for (int q = 0; q < NS; ++q) {
#pragma unroll
            for (int a = 0; a < S; ++a) acc[q][a] *= cur[q].get(a);
          }
        } break;
        case OP_STORE:
        case OP_ROOT: {
#pragma unroll
          for (int q = 0; q < NS; ++q) {
            double mx = acc[q][0];
#pragma unroll
            for (int a = 1; a < S; ++a) mx = fmax(mx, acc[q][a]);
            if (mx > 0.0) {
              const int e = rt_exponent(mx);
              const double sc = rt_pow2_neg(e);
#pragma unroll
              for (int a = 0; a < S; ++a) acc[q][a] *= sc;
              esum[q] += e;
            }
            if (STORE && op.w >= 0 && active[q]) {
              double* pp = partials + (int64_t)op.w * S * stride + site[q];
#pragma unroll
              for (int a = 0; a < S; ++a) __stcs(pp + (int64_t)a * stride, acc[q][a]);
              if (exponents) exponents[(int64_t)op.w * stride + site[q]] = esum[q];
            }
            if (op.x == OP_STORE) {
              double* sp = stk + op.z + q * S * kBlock + tid;
#pragma unroll
              for (int a = 0; a < S; ++a) sp[a * kBlock] = acc[q][a];
              estk[op.z / S + q * kBlock + tid] = esum[q];
#pragma unroll
              for (int a = 0; a < S; ++a) acc[q][a] = 1.0;
              esum[q] = 0;
            } else if (active[q]) {
              double lk = 0.0;
#pragma unroll
              for (int a = 0; a < S; ++a) lk = fma(pi_s[a], acc[q][a], lk);
              if (lk > 0.0) {
                const double ll = log(lk) + (double)esum[q] * RT_LN2;
                loglik[site[q]] = ll;
                status[site[q]] = RT_SITE_OK;
                my_ll += ll;
              } else {
                loglik[site[q]] = -INFINITY;
                status[site[q]] = RT_SITE_STRUCTURAL_ZERO;
              }
            }
          }
        } break;
        default: break;
      }
    }
  }

  if (loglik_sum) {
    __shared__ double red[kBlock / 32];
    double w = rt_warp_sum(my_ll);
    if ((tid & 31) == 0) red[tid >> 5] = w;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int i = 0; i < kBlock / 32; ++i) t += red[i];
      atomicAdd(loglik_sum, t);
    }
  }
}

template <int S, int OBS, bool STORE>
int launch_t(int64_t n_sites, int64_t stride, const int4* program, int n_ops, int n_slots,
             int n_nodes, const double* P, const double* root_distn, const void* obs,
             double* partials, int32_t* exponents, double* loglik, int8_t* status,
             double* loglik_sum, cudaStream_t stream, int packed = 0) {
  constexpr int NS = (S <= 4) ? 2 : 1;      // sites per thread
  size_t base = (size_t)n_ops * sizeof(int4) + sizeof(int) * (((size_t)n_ops + 4) & ~(size_t)3) +
                sizeof(double) * (S + (size_t)n_nodes * S);
  size_t stack = (size_t)n_slots * NS * S * kBlock * sizeof(double) +
                 (size_t)n_slots * NS * kBlock * sizeof(int) + 8;
  if (OBS == OBS_DENSE && kRing > 0)
    stack += (size_t)kRingSlots * NS * S * kBlock * sizeof(double) + sizeof(int) * ((size_t)n_ops + kRing + 2);
  size_t pbytes = (size_t)n_nodes * S * S * sizeof(double);
  const size_t limit = 200 * 1024;
  const bool p_in_smem = base + stack + pbytes <= 96 * 1024;
  size_t smem = base + stack + (p_in_smem ? pbytes : 0);
  if (smem > limit) return RT_ERR_UNSUPPORTED;
  const int64_t grid = (n_sites + (int64_t)kBlock * NS - 1) / ((int64_t)kBlock * NS);
  if (p_in_smem) {
    auto kern = prune_small_kernel<S, OBS, STORE, true, NS>;
    RT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, kBlock, smem, stream>>>(n_sites, stride, program, n_ops, n_slots, n_nodes,
                                                  P, root_distn, obs, partials, exponents, loglik,
                                                  status, loglik_sum, packed);
  } else {
    auto kern = prune_small_kernel<S, OBS, STORE, false, NS>;
    RT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, kBlock, smem, stream>>>(n_sites, stride, program, n_ops, n_slots, n_nodes,
                                                  P, root_distn, obs, partials, exponents, loglik,
                                                  status, loglik_sum, packed);
  }
  RT_CUDA_CHECK(cudaGetLastError());
  return RT_OK;
}

template <int S>
int launch_s(int obs_kind, bool store, int64_t n_sites, int64_t stride, const int4* program,
             int n_ops, int n_slots, int n_nodes, const double* P, const double* root_distn,
             const void* obs, double* partials, int32_t* exponents, double* loglik,
             int8_t* status, double* loglik_sum, cudaStream_t stream) {
#define RT_ARGS n_sites, stride, program, n_ops, n_slots, n_nodes, P, root_distn, obs, partials, \
                exponents, loglik, status, loglik_sum, stream
  switch (obs_kind) {
    case OBS_CODES: return store ? launch_t<S, OBS_CODES, true>(RT_ARGS) : launch_t<S, OBS_CODES, false>(RT_ARGS);
    case OBS_MASK:  return store ? launch_t<S, OBS_MASK, true>(RT_ARGS)  : launch_t<S, OBS_MASK, false>(RT_ARGS);
    case OBS_DENSE: return store ? launch_t<S, OBS_DENSE, true>(RT_ARGS) : launch_t<S, OBS_DENSE, false>(RT_ARGS);
    case 3:   // RT_OBS_CODES4: the codes kernels with the nibble decoder
      return store ? launch_t<S, OBS_CODES, true>(RT_ARGS, 1) : launch_t<S, OBS_CODES, false>(RT_ARGS, 1);
  }
#undef RT_ARGS
  return RT_ERR_ARG;
}

}  // namespace

int rt_prune_small_dispatch(int S, int obs_kind, bool store, int64_t n_sites, int64_t stride,
                            const int32_t* program, int n_ops, int n_slots, int n_nodes,
                            const double* P, const double* root_distn, const void* obs,
                            double* partials, int32_t* exponents, double* loglik, int8_t* status,
                            double* loglik_sum, cudaStream_t stream) {
  const int4* prog = reinterpret_cast<const int4*>(program);
#define RT_ARGS obs_kind, store, n_sites, stride, prog, n_ops, n_slots, n_nodes, P, root_distn, \
                obs, partials, exponents, loglik, status, loglik_sum, stream
  switch (S) {
    case 2: return launch_s<2>(RT_ARGS);
    case 3: return launch_s<3>(RT_ARGS);
    case 4: return launch_s<4>(RT_ARGS);
    case 5: return launch_s<5>(RT_ARGS);
    case 6: return launch_s<6>(RT_ARGS);
    case 7: return launch_s<7>(RT_ARGS);
    case 8: return launch_s<8>(RT_ARGS);
  }
#undef RT_ARGS
  return RT_ERR_UNSUPPORTED;
}
