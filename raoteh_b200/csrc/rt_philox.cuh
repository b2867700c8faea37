// Counter-based RNG shared by the Rao-Teh kernels.
#pragma once
#include <stdint.h>

// ---------------- Philox4x32-10 (Salmon et al. 2011), counter based ----------------
struct Philox {
  uint32_t key0, key1;
  uint32_t c0, c1, c2, c3;
  uint32_t out[4];
  int have;
  __device__ __forceinline__ void init(uint64_t seed, uint64_t traj, uint32_t sweep) {
    key0 = (uint32_t)seed; key1 = (uint32_t)(seed >> 32);
    c0 = 0; c1 = sweep; c2 = (uint32_t)traj; c3 = (uint32_t)(traj >> 32);
    have = 0;
  }
  __device__ __forceinline__ void round(uint32_t k0, uint32_t k1, uint32_t (&c)[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void refill() {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(k0, k1, c);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    ++c0;
    have = 4;
  }
  __device__ __forceinline__ uint32_t next() {
    if (have == 0) refill();
    --have;
    return out[have];
  }
  // counter-addressed block: 4 words for (substream, block) of this (trajectory, sweep);
  // every draw of the sweep has a fixed address, so lanes of a warp stay aligned inside
  // an edge's event loop and a lane's stream never depends on its neighbours
  __device__ __forceinline__ void block(uint32_t substream, uint32_t blk) {
    c0 = (substream << 16) | (blk & 0xffffu);
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(k0, k1, c);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
  // uniform in (0, 1]
  __device__ __forceinline__ float uniform() { return ((float)next() + 1.0f) * 2.3283064365386963e-10f; }
  __device__ __forceinline__ double uniform_d() { return ((double)next() + 0.5) * 2.3283064365386963e-10; }
};


#define RT_U32_TO_UNIT(u) (((float)(u) + 1.0f) * 2.3283064365386963e-10f)   /* (0, 1] */
