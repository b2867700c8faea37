// K4/K5 for small state spaces (S <= 8): downward pass, posterior node
// marginals and per-edge sufficient statistics, batched over sites.
//
// Replaces, batched over sites,
//   pyfelscore.mc0_esd_get_node_to_distn       (raoteh/sampler/_mc0_dense.py:381,
//        _mcy_dense.py:195; spec _mc0_dense.py:400-489)
//   pyfelscore.mc0_esd_get_joint_endpoint_distn (_mcy_dense.py:205; spec
//        _mc0_dense.py:217-270)
// and the `joint_prob / cond_prob` accumulation of
//   _mjp_dense.get_expected_history_statistics  (_mjp_dense.py:502-510,521-533).
//
// Edge-major and level-synchronous: one launch per tree level, grid =
// (site chunks) x (edges of the level).  A thread strides over the sites of its
// chunk, so the S*S weights W_b += G_b (x) L_b accumulate in registers with no
// cross-lane traffic until one block reduction + S*S atomics per CTA.  The
// joint J_b = G_b[:,None] * P_b * L_b[None,:] is never materialised.
// HBM traffic per site and edge: D_parent in, L_child in, D_child out
// (3 * S * 8 B; leaves read 1 code byte instead) -- HBM-bound.
#include "rt_common.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kSitesPerCta = 2048;

template <int S>
__global__ void __launch_bounds__(kBlock)
root_distn_kernel(int64_t n_sites, int64_t stride, const double* __restrict__ root_distn,
                  const double* __restrict__ partials, const int8_t* __restrict__ status,
                  double* __restrict__ node_distn, double* __restrict__ root_post_sum) {
  __shared__ double red[kBlock / 32][S];
  const int tid = threadIdx.x;
  double sum[S];
#pragma unroll
  for (int s = 0; s < S; ++s) sum[s] = 0.0;
  for (int64_t site = (int64_t)blockIdx.x * kBlock + tid; site < n_sites;
       site += (int64_t)gridDim.x * kBlock) {
    if (status[site] != RT_SITE_OK) {
#pragma unroll
      for (int s = 0; s < S; ++s) node_distn[(int64_t)s * stride + site] = 0.0;
      continue;
    }
    double w[S];
    double tot = 0.0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      w[s] = partials[(int64_t)s * stride + site] * (root_distn ? root_distn[s] : 1.0);
      tot += w[s];
    }
    const double inv = 1.0 / tot;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const double d = w[s] * inv;
      node_distn[(int64_t)s * stride + site] = d;
      sum[s] += d;
    }
  }
  if (root_post_sum) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      double v = rt_warp_sum(sum[s]);
      if ((tid & 31) == 0) red[tid >> 5][s] = v;
    }
    __syncthreads();
    if (tid < S) {
      double t = 0.0;
      for (int i = 0; i < kBlock / 32; ++i) t += red[i][tid];
      atomicAdd(&root_post_sum[tid], t);
    }
  }
}

template <int S, int OBS>
__global__ void __launch_bounds__(kBlock, (S <= 4 ? 2 : 1))
down_level_kernel(int64_t n_sites, int64_t stride, const int4* __restrict__ edges,
                  const double* __restrict__ P, const void* __restrict__ obs,
                  const double* __restrict__ partials, const int8_t* __restrict__ status,
                  double* __restrict__ node_distn, double* __restrict__ W) {
  __shared__ double Ps[S * S];
  __shared__ double red[kBlock / 32][S * S];
  const int tid = threadIdx.x;
  const int4 e = edges[blockIdx.y];   // (child node, parent store, child store, child obs slot)
  const int b = e.x;
  if (tid < S * S) Ps[tid] = P[(size_t)b * S * S + tid];
  __syncthreads();

  double Wacc[S][S];
#pragma unroll
  for (int i = 0; i < S; ++i)
#pragma unroll
    for (int j = 0; j < S; ++j) Wacc[i][j] = 0.0;

  const double* Dp = node_distn + (int64_t)e.y * S * stride;
  const double* Lc = e.z >= 0 ? partials + (int64_t)e.z * S * stride : nullptr;
  double* Dc = e.z >= 0 ? node_distn + (int64_t)e.z * S * stride : nullptr;

  const int64_t lo = (int64_t)blockIdx.x * kSitesPerCta;
  const int64_t hi = (n_sites < lo + kSitesPerCta) ? n_sites : lo + kSitesPerCta;

  // loads of one site (software-pipelined: the next site's loads are in flight
  // while the current one is computed)
  auto load_site = [&](int64_t site, double (&L)[S], double (&D)[S], bool& ok) {
    ok = site < hi && status[site] == RT_SITE_OK;
    if (!ok) {
#pragma unroll
      for (int s = 0; s < S; ++s) { L[s] = 0.0; D[s] = 0.0; }
      return;
    }
    if (Lc) {
#pragma unroll
      for (int s = 0; s < S; ++s) L[s] = __ldcs(&Lc[(int64_t)s * stride + site]);
    } else if (e.w < 0) {
#pragma unroll
      for (int s = 0; s < S; ++s) L[s] = 1.0;
    } else if (OBS == OBS_CODES) {
      const int k = reinterpret_cast<const uint8_t*>(obs)[(int64_t)e.w * stride + site];
#pragma unroll
      for (int s = 0; s < S; ++s) L[s] = (k == RT_MISSING || k == s) ? 1.0 : 0.0;
    } else if (OBS == OBS_MASK) {
      const unsigned long long mk =
          reinterpret_cast<const unsigned long long*>(obs)[(int64_t)e.w * stride + site];
#pragma unroll
      for (int s = 0; s < S; ++s) L[s] = ((mk >> s) & 1ull) ? 1.0 : 0.0;
    } else {
      const double* d = reinterpret_cast<const double*>(obs);
#pragma unroll
      for (int s = 0; s < S; ++s) L[s] = d[((int64_t)e.w * S + s) * stride + site];
    }
#pragma unroll
    for (int s = 0; s < S; ++s) D[s] = Dp[(int64_t)s * stride + site];
  };

  double Ln[S], Dn[S];
  bool okn;
  load_site(lo + tid, Ln, Dn, okn);
  for (int64_t site = lo + tid; site < hi; site += kBlock) {
    double L[S], D[S];
    const bool ok = okn;
#pragma unroll
    for (int s = 0; s < S; ++s) { L[s] = Ln[s]; D[s] = Dn[s]; }
    load_site(site + kBlock, Ln, Dn, okn);
    if (!ok) {
      if (Dc) {
#pragma unroll
        for (int s = 0; s < S; ++s) Dc[(int64_t)s * stride + site] = 0.0;
      }
      continue;
    }
    double G[S];
#pragma unroll
    for (int a = 0; a < S; ++a) {
      double m = 0.0;
#pragma unroll
      for (int c = 0; c < S; ++c) m = fma(Ps[a * S + c], L[c], m);
      G[a] = (D[a] > 0.0) ? D[a] / m : 0.0;
    }
    if (Dc) {
#pragma unroll
      for (int c = 0; c < S; ++c) {
        double t = 0.0;
#pragma unroll
        for (int a = 0; a < S; ++a) t = fma(G[a], Ps[a * S + c], t);
        Dc[(int64_t)c * stride + site] = t * L[c];
      }
    }
#pragma unroll
    for (int a = 0; a < S; ++a)
#pragma unroll
      for (int c = 0; c < S; ++c) Wacc[a][c] = fma(G[a], L[c], Wacc[a][c]);
  }

#pragma unroll
  for (int a = 0; a < S; ++a)
#pragma unroll
    for (int c = 0; c < S; ++c) {
      double v = rt_warp_sum(Wacc[a][c]);
      if ((tid & 31) == 0) red[tid >> 5][a * S + c] = v;
    }
  __syncthreads();
  if (tid < S * S) {
    double t = 0.0;
    for (int i = 0; i < kBlock / 32; ++i) t += red[i][tid];
    if (Ps[tid] > 0.0 && t != 0.0) atomicAdd(&W[(size_t)b * S * S + tid], t);
  }
}

template <int S>
int run(int obs_kind, int64_t n_sites, int64_t stride, const int32_t* edges_dev,
        const int32_t* level_ptr_h, int n_levels, const double* P, const double* root_distn,
        const void* obs, const double* partials, const int8_t* status, double* node_distn,
        double* W, double* root_post_sum, cudaStream_t stream) {
  int64_t gr = (n_sites + kBlock - 1) / kBlock;
  int grid_root = (int)(gr < 148 * 8 ? gr : 148 * 8);
  root_distn_kernel<S><<<grid_root, kBlock, 0, stream>>>(n_sites, stride, root_distn, partials,
                                                         status, node_distn, root_post_sum);
  const int4* edges = reinterpret_cast<const int4*>(edges_dev);
  const unsigned gx = (unsigned)((n_sites + kSitesPerCta - 1) / kSitesPerCta);
  for (int l = 0; l < n_levels; ++l) {
    const int e0 = level_ptr_h[l], e1 = level_ptr_h[l + 1];
    if (e1 <= e0) continue;
    dim3 grid(gx, (unsigned)(e1 - e0));
    switch (obs_kind) {
      case OBS_CODES:
        down_level_kernel<S, OBS_CODES><<<grid, kBlock, 0, stream>>>(
            n_sites, stride, edges + e0, P, obs, partials, status, node_distn, W);
        break;
      case OBS_MASK:
        down_level_kernel<S, OBS_MASK><<<grid, kBlock, 0, stream>>>(
            n_sites, stride, edges + e0, P, obs, partials, status, node_distn, W);
        break;
      case OBS_DENSE:
        down_level_kernel<S, OBS_DENSE><<<grid, kBlock, 0, stream>>>(
            n_sites, stride, edges + e0, P, obs, partials, status, node_distn, W);
        break;
      default: return RT_ERR_ARG;
    }
  }
  RT_CUDA_CHECK(cudaGetLastError());
  return RT_OK;
}

}  // namespace

int rt_posterior_small_dispatch(int S, int obs_kind, int64_t n_sites, int64_t stride,
                                const int32_t* edges_dev, const int32_t* level_ptr_h, int n_levels,
                                const double* P, const double* root_distn, const void* obs,
                                const double* partials, const int8_t* status, double* node_distn,
                                double* W, double* root_post_sum, cudaStream_t stream) {
#define RT_ARGS obs_kind, n_sites, stride, edges_dev, level_ptr_h, n_levels, P, root_distn, obs, \
                partials, status, node_distn, W, root_post_sum, stream
  switch (S) {
    case 2: return run<2>(RT_ARGS);
    case 3: return run<3>(RT_ARGS);
    case 4: return run<4>(RT_ARGS);
    case 5: return run<5>(RT_ARGS);
    case 6: return run<6>(RT_ARGS);
    case 7: return run<7>(RT_ARGS);
    case 8: return run<8>(RT_ARGS);
  }
#undef RT_ARGS
  return RT_ERR_UNSUPPORTED;
}
