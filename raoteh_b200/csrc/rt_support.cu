// A4: structural support sets, batched over sites (integer / bit work).
//
// Replaces pyfelscore.mcy_esd_get_node_to_pset + esd_get_node_to_set
// (raoteh/sampler/_mcy_dense.py:168-179,270-281; _mcy.py:219-230,508-519;
// specs _mcy.py:397-470 and _mc0.py:89-138).  One thread per site walks the
// tree twice over uint64 state bitmasks; the P > 0 pattern of every edge is
// pre-packed into one uint64 row mask per (edge, state).
#include "rt_common.cuh"

namespace {

__global__ void pack_pattern_kernel(const double* __restrict__ P, int S, int n_nodes,
                                    unsigned long long* __restrict__ rowbits) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (node, s)
  if (idx >= n_nodes * S) return;
  const double* row = P + (size_t)idx * S;
  unsigned long long m = 0ull;
  for (int c = 0; c < S; ++c)
    if (row[c] > 0.0) m |= (1ull << c);
  rowbits[idx] = m;
}

__global__ void __launch_bounds__(128)
support_kernel(int S, int n_nodes, int64_t n_sites, int64_t stride, int passes,
               const int32_t* __restrict__ parent,
               const unsigned long long* __restrict__ rowbits,
               unsigned long long* __restrict__ mask) {
  const int64_t site = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= n_sites) return;
  // backward: state s of a kept iff every child b has some kept s' with P_b[s,s'] > 0
  for (int b = n_nodes - 1; b >= 1 && (passes & 1); --b) {
    const int a = parent[b];
    const unsigned long long Mb = mask[(int64_t)b * stride + site];
    unsigned long long keep = 0ull;
    const unsigned long long* rb = rowbits + (size_t)b * S;
    for (int s = 0; s < S; ++s)
      if (rb[s] & Mb) keep |= (1ull << s);
    mask[(int64_t)a * stride + site] &= keep;
  }
  // forward: state s' of b kept iff reachable from a kept state of its parent
  for (int b = 1; b < n_nodes && (passes & 2); ++b) {
    const int a = parent[b];
    unsigned long long Ma = mask[(int64_t)a * stride + site];
    unsigned long long reach = 0ull;
    const unsigned long long* rb = rowbits + (size_t)b * S;
    while (Ma) {
      const int s = __ffsll((long long)Ma) - 1;
      Ma &= Ma - 1;
      reach |= rb[s];
    }
    mask[(int64_t)b * stride + site] &= reach;
  }
}

}  // namespace

int rt_support_sets_impl(int S, int n_nodes, int64_t n_sites, int64_t stride, int passes, const int32_t* parent,
                         const double* P, uint64_t* mask, cudaStream_t stream) {
  if (S < 1 || S > 64) return RT_ERR_UNSUPPORTED;
  if (n_nodes <= 0 || n_sites <= 0) return RT_OK;
  unsigned long long* rowbits = nullptr;
  RT_CUDA_CHECK(rt_ws_alloc((void**)&rowbits, sizeof(unsigned long long) * (size_t)n_nodes * S, stream));
  const int tot = n_nodes * S;
  pack_pattern_kernel<<<(tot + 255) / 256, 256, 0, stream>>>(P, S, n_nodes, rowbits);
  support_kernel<<<(unsigned)((n_sites + 127) / 128), 128, 0, stream>>>(
      S, n_nodes, n_sites, stride, passes, parent, rowbits,
      reinterpret_cast<unsigned long long*>(mask));
  cudaError_t e = cudaGetLastError();
  rt_ws_free(rowbits, stream);
  RT_CUDA_CHECK(e);
  return RT_OK;
}
