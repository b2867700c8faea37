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
#include "rt_small.cuh"

namespace {

constexpr int kBlock = 128;

using rt_small::ObsVal;
using rt_small::max_hiword;

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
  long long* pre_s = reinterpret_cast<long long*>(prog_s + n_ops);     // [n_ops + 1], padded to even
  double* pi_s = reinterpret_cast<double*>(pre_s + ((n_ops + 2) & ~1));
  double* zero_s = pi_s + S;                     // [2], zeros
  double* rowsum_s = zero_s + 2;
  double* P_s = rowsum_s + n_nodes * S;          // 16-byte aligned for even S
  double* stk = P_s + (P_SMEM ? n_nodes * S * S : 0);
  int* estk = reinterpret_cast<int*>(stk + n_slots * NS * S * kBlock);

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
  if constexpr (OBS == OBS_CODES) {
    // every code row of this CTA's sites into L2 now, while the tables below are set up: a leaf
    // op is four multiplies, so the one-op-ahead fetch of its code byte cannot hide a DRAM round
    // trip (19 % of the warp samples sat on the first use of the byte, ncu source view)
    const int64_t s0 = (int64_t)blockIdx.x * NS * kBlock;
    const int64_t b0 = obs_packed ? (s0 >> 1) : s0;
    const int64_t nb = obs_packed ? (NS * kBlock / 2) : (NS * kBlock);
    const int lines = (int)((nb + 127) / 128) + 1;          // + 1: rows are not line aligned
    const int64_t row_bytes = obs_packed ? ((stride + 1) >> 1) : stride;
    for (int i = tid; i < n_ops * lines; i += kBlock) {
      const int4 d = prog_s[i / lines];
      if (d.x == OP_MSG_OBS || d.x == OP_APPLY_OBS) {
        int64_t off = b0 + (int64_t)(i % lines) * 128;
        if (off >= row_bytes) off = row_bytes - 1;
        asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const uint8_t*>(obs) +
                     (int64_t)d.z * row_bytes + off));
      }
    }
  }
  if (tid == 0) {
    // pre_s[ip]: element offset of the row consumed by the next obs-consuming op after ip
    long long nxt = -1;
    for (int i = n_ops - 1; i >= 0; --i) {
      pre_s[i] = nxt;
      const int code = prog_s[i].x;
      if (code == OP_MSG_OBS || code == OP_APPLY_OBS) nxt = ObsVal<S, OBS>::row_off(prog_s[i].z, stride, obs_packed);
    }
    pre_s[n_ops] = nxt;     // first row of the program
  }
  if (tid < S) pi_s[tid] = root_distn ? root_distn[tid] : 1.0;
  if (tid < 2) zero_s[tid] = 0.0;
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
    // cur[q] holds the row of the NEXT obs-consuming op: an obs op computes from it and then
    // fetches its successor into the same registers (no second buffer, no copies)
    ObsVal<S, OBS> cur[NS];
    const void* obs_q[NS];
    int nib[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
#pragma unroll
      for (int a = 0; a < S; ++a) acc[q][a] = 1.0;
      esum[q] = 0;
      obs_q[q] = ObsVal<S, OBS>::base(obs, site[q], obs_packed);
      nib[q] = (int)(site[q] & 1) * 4;
      cur[q].init();
      cur[q].fetch(obs_q[q], pre_s[n_ops], stride);
    }

    for (int ip = 0; ip < n_ops; ++ip) {
      const int4 op = prog_s[ip];
      const double* Pc = P_SMEM ? (P_s + op.y) : (P + op.y);
      // m[q][a] = sum_b P[a][b] v[q][b]; even S with P in shared memory: 16-byte loads
      auto matvec = [&](const double (&v)[NS][S], double (&m)[NS][S]) {
#pragma unroll
        for (int a = 0; a < S; ++a) {
#pragma unroll
          for (int q = 0; q < NS; ++q) m[q][a] = 0.0;
          if constexpr (P_SMEM && (S % 2 == 0)) {
            const double2* pr = reinterpret_cast<const double2*>(Pc + a * S);
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
              const double p = Pc[a * S + b];
#pragma unroll
              for (int q = 0; q < NS; ++q) m[q][a] = fma(p, v[q][b], m[q][a]);
            }
          }
        }
      };
      switch (op.x) {
        case OP_MSG_SLOT: {
          double v[NS][S], m[NS][S];
#pragma unroll
          for (int q = 0; q < NS; ++q) {
            const double* sp = stk + op.z + q * S * kBlock + tid;
#pragma unroll
            for (int b = 0; b < S; ++b) v[q][b] = sp[b * kBlock];
            esum[q] += estk[op.z / S + q * kBlock + tid];
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
              const int k = cur[q].code(obs_packed ? nib[q] : -1);
              if constexpr (P_SMEM) {
                // branch-free: column k of P, the row sums, or zeros (invalid code), all in
                // shared memory (divergent writes to acc made the compiler copy acc every op)
                const double* src = (k == RT_MISSING) ? (rowsum_s + op.y / S) : (k < S ? Pc + k : zero_s);
                const int strd = (k == RT_MISSING) ? 1 : (k < S ? S : 0);
#pragma unroll
                for (int a = 0; a < S; ++a) acc[q][a] *= src[a * strd];
              } else if (k == RT_MISSING) {
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
              for (int b = 0; b < S; ++b) v[q][b] = cur[q].get(b, obs_packed ? nib[q] : -1);
            matvec(v, m);
#pragma unroll
            for (int q = 0; q < NS; ++q)
#pragma unroll
              for (int a = 0; a < S; ++a) acc[q][a] *= m[q][a];
          }
          const long long off = pre_s[ip];
#pragma unroll
          for (int q = 0; q < NS; ++q) cur[q].fetch(obs_q[q], off, stride);
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
            for (int a = 0; a < S; ++a) acc[q][a] *= cur[q].get(a, obs_packed ? nib[q] : -1);
          }
          const long long off = pre_s[ip];
#pragma unroll
          for (int q = 0; q < NS; ++q) cur[q].fetch(obs_q[q], off, stride);
        } break;
        case OP_STORE:
        case OP_ROOT: {
#pragma unroll
          for (int q = 0; q < NS; ++q) {
            const int hmx = max_hiword<S>(acc[q]);
            if (hmx >= 0x00100000) {          // largest entry is a positive normal number
              const int e = (hmx >> 20) - 1023;
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
  size_t base = (size_t)n_ops * sizeof(int4) + sizeof(long long) * (((size_t)n_ops + 2) & ~(size_t)1) +
                sizeof(double) * (S + 2 + (size_t)n_nodes * S);
  size_t stack = (size_t)n_slots * NS * S * kBlock * sizeof(double) +
                 (size_t)n_slots * NS * kBlock * sizeof(int) + 8;
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
