// Shared definitions for the raoteh_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

// ---- status codes returned by every C-ABI entry point ----------------------
#define RT_OK 0
#define RT_ERR_ARG 1        // bad argument (size, null pointer, unsupported S)
#define RT_ERR_CUDA 2       // CUDA runtime error (see rt_last_error_string)
#define RT_ERR_UNSUPPORTED 3

// ---- per-site status written by the kernels (reference: _util.py:14-21) ----
#define RT_SITE_OK 0
#define RT_SITE_STRUCTURAL_ZERO 1   // StructuralZeroProb
#define RT_SITE_NUMERICAL_ZERO 2    // NumericalZeroProb

// ---- upward-program op codes (mirrors raoteh_b200/lowering.py) --------------
enum {
  OP_MSG_SLOT = 0,
  OP_MSG_OBS = 1,
  OP_MSG_ONES = 2,
  OP_APPLY_OBS = 3,
  OP_STORE = 4,
  OP_ROOT = 5
};
// flags in bits 8.. of the op code (raoteh_b200/lowering.py)
#define OP_FLAG_FRESH (1 << 8)          // OP_MSG_SLOT: the child's partial was stored by the previous op
#define OP_FLAG_KEEP_ON_CHIP (1 << 9)   // OP_STORE: the next op consumes this partial FRESH
#define OP_FLAG_PARK (1 << 10)          // OP_STORE: a later, non-fresh op reads it back from its slot

// ---- observation encodings ---------------------------------------------------
enum {
  OBS_CODES = 0,   // uint8  [n_obs][stride]        hard state, 255 = unobserved
  OBS_MASK = 1,    // uint64 [n_obs][stride]        bitmask of allowed states
  OBS_DENSE = 2    // double [n_obs][S][stride]     emission likelihoods
};
#define RT_MISSING 255

#define RT_LN2 0.69314718055994530942

#define RT_CUDA_CHECK(expr)                         \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) {                        \
      rt_set_last_error(_e, __FILE__, __LINE__);    \
      return RT_ERR_CUDA;                           \
    }                                               \
  } while (0)

void rt_set_last_error(cudaError_t e, const char* file, int line);

// Per-call workspaces come from a PRIVATE stream-ordered pool of the current device (rt_api.cu),
// not from the device's default pool: freed blocks stay cached in it between calls (no allocator
// round trip per launch) without changing the behaviour of anybody else's cudaMallocAsync.
// rt_release_workspace() (C ABI) trims it.
cudaError_t rt_ws_alloc(void** p, size_t bytes, cudaStream_t stream);
cudaError_t rt_ws_free(void* p, cudaStream_t stream);

// exponent of a positive finite double (floor(log2 x)); 0 for x == 0
__device__ __forceinline__ int rt_exponent(double x) {
  int hi = __double2hiint(x);
  return ((hi >> 20) & 0x7ff) - 1023;
}
// 2^(-e) for |e| <= 1022, built from bits (exact)
__device__ __forceinline__ double rt_pow2_neg(int e) {
  return __hiloint2double((1023 - e) << 20, 0);
}

__device__ __forceinline__ double rt_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
