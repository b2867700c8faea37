// A8 in materialised form: per-edge posterior joint distributions and the
// posterior marginal of EVERY node (leaves included), for small batches.
//
// Replaces pyfelscore.mc0_esd_get_joint_endpoint_distn (raoteh/sampler/
// _mcy_dense.py:205; spec _mc0_dense.py:217-270) and the leaf rows of
// mc0_esd_get_node_to_distn (_mc0_dense.py:381).  The batched hot path never
// materialises J (rt_posterior_stats consumes it on chip); this kernel exists
// for the reference-shaped API (`get_joint_endpoint_distn`, `kitchen_sink`),
// whose return value IS the joint.  One CTA per (edge, site).
#include "rt_common.cuh"

namespace {

template <int OBS>
__global__ void __launch_bounds__(64)
joint_kernel(int S, int64_t n_sites, int64_t stride, const int4* __restrict__ edges,
             const double* __restrict__ P, const void* __restrict__ obs,
             const double* __restrict__ partials, const double* __restrict__ node_distn,
             const int8_t* __restrict__ status, double* __restrict__ J, double* __restrict__ D_all) {
  __shared__ double L[64], G[64];
  const int4 e = edges[blockIdx.x];   // (child node, parent store, child store, child obs slot)
  const int64_t site = blockIdx.y;
  const int b = e.x;
  const int tid = threadIdx.x;
  const double* Pb = P + (size_t)b * S * S;
  double* Jb = J + ((size_t)b * n_sites + site) * S * S;
  double* Db = D_all + ((size_t)b * n_sites + site) * S;
  const bool ok = status[site] == RT_SITE_OK;
  if (tid < S) {
    double l;
    if (e.z >= 0) l = partials[((int64_t)e.z * S + tid) * stride + site];
    else if (e.w < 0) l = 1.0;
    else if (OBS == OBS_CODES) {
      const int k = reinterpret_cast<const uint8_t*>(obs)[(int64_t)e.w * stride + site];
      l = (k == RT_MISSING || k == tid) ? 1.0 : 0.0;
    } else if (OBS == OBS_MASK) {
      const unsigned long long mk =
          reinterpret_cast<const unsigned long long*>(obs)[(int64_t)e.w * stride + site];
      l = ((mk >> tid) & 1ull) ? 1.0 : 0.0;
    } else {
      l = reinterpret_cast<const double*>(obs)[((int64_t)e.w * S + tid) * stride + site];
    }
    L[tid] = l;
  }
  __syncthreads();
  if (tid < S) {
    double m = 0.0;
    for (int c = 0; c < S; ++c) m = fma(Pb[tid * S + c], L[c], m);
    const double d = ok ? node_distn[((int64_t)e.y * S + tid) * stride + site] : 0.0;
    G[tid] = (d > 0.0 && m > 0.0) ? d / m : 0.0;
  }
  __syncthreads();
  if (tid < S) {
    // column tid of J, and its sum = posterior marginal of the child
    double col = 0.0;
    for (int a = 0; a < S; ++a) {
      const double v = G[a] * Pb[a * S + tid] * L[tid];
      Jb[a * S + tid] = v;
      col += v;
    }
    Db[tid] = col;
  }
}

}  // namespace

int rt_joint_distn_impl(int S, int obs_kind, int64_t n_sites, int64_t stride,
                        const int32_t* edges, int n_edges, const double* P, const void* obs,
                        const double* partials, const double* node_distn, const int8_t* status,
                        double* J, double* D_all, cudaStream_t stream) {
  if (S < 1 || S > 64) return RT_ERR_UNSUPPORTED;
  if (n_edges <= 0 || n_sites <= 0) return RT_OK;
  if (n_sites > 65535) return RT_ERR_ARG;
  const int4* ed = reinterpret_cast<const int4*>(edges);
  dim3 grid((unsigned)n_edges, (unsigned)n_sites);
  switch (obs_kind) {
    case OBS_CODES:
      joint_kernel<OBS_CODES><<<grid, 64, 0, stream>>>(S, n_sites, stride, ed, P, obs, partials,
                                                       node_distn, status, J, D_all); break;
    case OBS_MASK:
      joint_kernel<OBS_MASK><<<grid, 64, 0, stream>>>(S, n_sites, stride, ed, P, obs, partials,
                                                      node_distn, status, J, D_all); break;
    case OBS_DENSE:
      joint_kernel<OBS_DENSE><<<grid, 64, 0, stream>>>(S, n_sites, stride, ed, P, obs, partials,
                                                       node_distn, status, J, D_all); break;
    default: return RT_ERR_ARG;
  }
  RT_CUDA_CHECK(cudaGetLastError());
  return RT_OK;
}
