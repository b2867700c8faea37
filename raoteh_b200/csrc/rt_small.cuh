// Shared pieces of the small-state-space kernels (S <= 8): observation rows fetched one op ahead,
// the integer-max exponent, the branch-free reciprocal.
#pragma once
#include "rt_common.cuh"

namespace rt_small {

// One observation row of one site, fetched one obs-consuming op ahead of its use
// (one specialisation per encoding so that only the live member occupies registers).
// The row's element offset is precomputed per op in shared memory (row_off) and the thread
// keeps one base pointer per site (base), so a fetch is a 64-bit add plus the loads.
template <int S, int OBS> struct ObsVal;

template <int S> struct ObsVal<S, OBS_CODES> {
  int k;
  __device__ __forceinline__ void init() { k = RT_MISSING; }
  // packed: two codes per byte (RT_OBS_CODES4), site i in nibble i & 1 of byte i >> 1, 15 = unobserved
  static __device__ __forceinline__ long long row_off(int row, int64_t stride, int packed) {
    return (long long)row * (packed ? ((stride + 1) >> 1) : stride);
  }
  static __device__ __forceinline__ const void* base(const void* obs, int64_t site, int packed) {
    return reinterpret_cast<const uint8_t*>(obs) + (packed ? (site >> 1) : site);
  }
  // the raw byte is kept and decoded at its use one op later (decoding here would make the
  // thread wait for the load it has just issued)
  __device__ __forceinline__ void fetch(const void* __restrict__ tb, long long off, int64_t) {
    if (off >= 0) k = reinterpret_cast<const uint8_t*>(tb)[off];
  }
  // nib < 0: one code per byte; else the shift of this site's nibble (byte 255 = both unobserved)
  __device__ __forceinline__ int code(int nib) const {
    if (nib < 0) return k;
    const int c = (k >> nib) & 15;
    return c == 15 ? RT_MISSING : c;
  }
  __device__ __forceinline__ double get(int b, int nib) const {
    const int c = code(nib);
    return (c == RT_MISSING || c == b) ? 1.0 : 0.0;
  }
};

template <int S> struct ObsVal<S, OBS_MASK> {
  unsigned long long mk;
  __device__ __forceinline__ void init() { mk = ~0ull; }
  static __device__ __forceinline__ long long row_off(int row, int64_t stride, int) {
    return (long long)row * stride;
  }
  static __device__ __forceinline__ const void* base(const void* obs, int64_t site, int) {
    return reinterpret_cast<const unsigned long long*>(obs) + site;
  }
  __device__ __forceinline__ void fetch(const void* __restrict__ tb, long long off, int64_t) {
    if (off >= 0) mk = reinterpret_cast<const unsigned long long*>(tb)[off];
  }
  __device__ __forceinline__ double get(int b, int) const { return ((mk >> b) & 1ull) ? 1.0 : 0.0; }
};

template <int S> struct ObsVal<S, OBS_DENSE> {
  double d[S];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < S; ++s) d[s] = 1.0;
  }
  static __device__ __forceinline__ long long row_off(int row, int64_t stride, int) {
    return (long long)row * S * stride;
  }
  static __device__ __forceinline__ const void* base(const void* obs, int64_t site, int) {
    return reinterpret_cast<const double*>(obs) + site;
  }
  __device__ __forceinline__ void fetch(const void* __restrict__ tb, long long off, int64_t stride) {
    if (off < 0) return;
    const double* p = reinterpret_cast<const double*>(tb) + off;
#pragma unroll
    for (int s = 0; s < S; ++s) d[s] = __ldcs(p + (int64_t)s * stride);
  }
  __device__ __forceinline__ double get(int b, int) const { return d[b]; }
};

// Recorded experiment (removed from the source, numbers in profiles/r1_dense_ring_sweep.log):
// dense emission rows streamed through a per-thread cp.async ring in shared memory were
// SLOWER on B200 than the one-row register prefetch (C2, 1e6 sites: ring 0/1/2/4 rows ->
// 0.287/0.360/0.392/0.524 ms): the ring's shared memory costs more occupancy than the extra
// bytes in flight buy, and the kernel's floor is its instruction issue.

// exponent field of the largest of S non-negative doubles via an integer max of their high
// words (one VIMNMX per state instead of the ~7-instruction IEEE fmax); 0 for zeros/subnormals
template <int S>
__device__ __forceinline__ int max_hiword(const double (&a)[S]) {
  int h = __double2hiint(a[0]);
#pragma unroll
  for (int s = 1; s < S; ++s) h = max(h, __double2hiint(a[s]));
  return h;
}


// 1/x to ~1 ulp: hardware seed (rcp.approx.ftz.f64, ~20 bits, full double range) plus two
// Newton steps.  Replaces the ~20-instruction IEEE division; the quotient is within 2 ulp,
// far inside the 1e-10 tolerance of the path.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * fma(-x, r, 2.0);
  r = r * fma(-x, r, 2.0);
  return r;
}

}  // namespace rt_small
