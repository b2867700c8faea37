// K1: batched matrix exponential, one CTA per matrix, Pade-13 scaling and
// squaring (Higham 2005), fp64.
//
// Replaces the reference's per-edge-per-site scipy.linalg.expm
// (raoteh/sampler/_mjp_dense.py:24,357; sparse twin _linalg.py:72-90) with one
// launch for every edge of the tree, computed once per (Q, tree) and shared by
// all sites.  The same kernel exponentiates the 2S x 2S block-triangular
// matrix [[tQ^T, tW],[0, tQ^T]] whose top-right block is the Frechet
// derivative L(tQ^T, tW) that contracts the posterior edge weights into
// expected dwell times / transition counts (replaces the S + nnz(Q)
// scipy.linalg.expm_frechet calls per edge per site at _mjp_dense.py:497-520).
//
// Work is negligible next to pruning, so matrices
// live in a global scratch (L2 resident) and products are smem-tiled FMA.
#include "rt_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 64;   // C tile 64x64, 4x4 per thread
constexpr int kTK = 16;

__device__ const double kPade13[14] = {
    64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0,
    129060195264000.0,   10559470521600.0,    670442572800.0,     33522128640.0,
    1323241920.0,        40840800.0,          960960.0,           16380.0,
    182.0,               1.0};
constexpr double kTheta13 = 5.371920351148152;

// C = A * B, all n x n row-major in global memory; whole CTA cooperates.
__device__ void mm(double* __restrict__ C, const double* __restrict__ A,
                   const double* __restrict__ B, int n, double* sm) {
  double (*As)[kTK + 1] = reinterpret_cast<double (*)[kTK + 1]>(sm);             // [64][17]
  double (*Bs)[kTile + 4] = reinterpret_cast<double (*)[kTile + 4]>(sm + kTile * (kTK + 1));  // [16][68]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, each 4x4
  for (int i0 = 0; i0 < n; i0 += kTile) {
    for (int j0 = 0; j0 < n; j0 += kTile) {
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
      for (int k0 = 0; k0 < n; k0 += kTK) {
        __syncthreads();
        for (int idx = tid; idx < kTile * kTK; idx += kThreads) {
          int r = idx / kTK, c = idx % kTK;
          int gi = i0 + r, gk = k0 + c;
          As[r][c] = (gi < n && gk < n) ? A[(size_t)gi * n + gk] : 0.0;
        }
        for (int idx = tid; idx < kTK * kTile; idx += kThreads) {
          int r = idx / kTile, c = idx % kTile;
          int gk = k0 + r, gj = j0 + c;
          Bs[r][c] = (gk < n && gj < n) ? B[(size_t)gk * n + gj] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTK; ++k) {
          double a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][k];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];     // lanes on consecutive columns: no bank conflicts
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int gi = i0 + ty * 4 + i, gj = j0 + tx + 16 * j;
          if (gi < n && gj < n) C[(size_t)gi * n + gj] = acc[i][j];
        }
    }
  }
  __syncthreads();
}

__device__ double block_max(double v, double* red) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int i = 1; i < kThreads / 32; ++i) r = fmax(r, red[i]);
  __syncthreads();
  return r;
}

// In-place expm of mats[blockIdx.x] (n x n).  Six n x n work matrices in `scratch` (7 n^2 doubles
// per matrix are reserved there); with SMEM_WS (n <= 64 or so) the three that the elimination and
// the squarings work on are in dynamic shared memory instead -- the elimination is 61 dependent
// steps of short loads, which took most of the 235 us per matrix when they went to L2.
template <bool SMEM_WS>
__global__ void __launch_bounds__(kThreads)
expm_kernel(double* __restrict__ mats, int n, double* __restrict__ scratch) {
  __shared__ double sm[kTile * (kTK + 1) + kTK * (kTile + 4)];
  __shared__ double red[kThreads / 32];
  __shared__ int piv_s;
  extern __shared__ double ws_dyn[];
  const int tid = threadIdx.x;
  const size_t nn = (size_t)n * n;
  double* A = mats + (size_t)blockIdx.x * nn;
  // the powers (read through the tile staging of mm and by streaming loops only) stay in the
  // global scratch; the three matrices the elimination and the squarings work on are in shared
  // memory: 3 * 30 KB at n = 61, so two CTAs share an SM and 255 matrices are one wave
  double* G = scratch + (size_t)blockIdx.x * 7 * nn;
  double *A2 = G, *A4 = G + nn, *A6 = G + 2 * nn;
  double* W = SMEM_WS ? ws_dyn : G + 3 * nn;
  double *T1 = W, *U = W + nn, *V = W + 2 * nn;

  // 1-norm (max column sum of |a_ij|)
  double cmax = 0.0;
  for (int j = tid; j < n; j += kThreads) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += fabs(A[(size_t)i * n + j]);
    cmax = fmax(cmax, s);
  }
  const double norm1 = block_max(cmax, red);
  int s = 0;
  if (norm1 > kTheta13) {
    s = (int)ceil(log2(norm1 / kTheta13));
    if (s < 0) s = 0;
    const double sc = ldexp(1.0, -s);
    for (size_t i = tid; i < nn; i += kThreads) A[i] *= sc;
  }
  __syncthreads();

  const double* b = kPade13;
  mm(A2, A, A, n, sm);
  mm(A4, A2, A2, n, sm);
  mm(A6, A4, A2, n, sm);
  // U = A6 (b13 A6 + b11 A4 + b9 A2), V = A6 (b12 A6 + b10 A4 + b8 A2), one temporary for both
  for (size_t i = tid; i < nn; i += kThreads) T1[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i];
  __syncthreads();
  mm(U, A6, T1, n, sm);
  for (size_t i = tid; i < nn; i += kThreads) T1[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i];
  __syncthreads();
  mm(V, A6, T1, n, sm);
  for (size_t i = tid; i < nn; i += kThreads) {
    const double a2 = A2[i], a4 = A4[i], a6 = A6[i];
    const bool diag = (i / n) == (i % n);
    T1[i] = U[i] + b[7] * a6 + b[5] * a4 + b[3] * a2 + (diag ? b[1] : 0.0);
    V[i] = V[i] + b[6] * a6 + b[4] * a4 + b[2] * a2 + (diag ? b[0] : 0.0);
  }
  __syncthreads();
  mm(U, A, T1, n, sm);      // U = A * (...)
  // M = V - U (into T1, dead after the product), B = V + U (in place)
  double* M = T1;
  double* B = V;
  for (size_t i = tid; i < nn; i += kThreads) {
    const double u = U[i], v = V[i];
    M[i] = v - u;
    B[i] = v + u;
  }
  __syncthreads();

  // Solve M X = B by Gauss-Jordan elimination with partial pivoting (every other row is
  // eliminated in step k, so there is no serial back substitution); X -> B.
  for (int k = 0; k < n; ++k) {
    if (tid < 32) {
      double best = -1.0;
      int bi = k;
      for (int i = k + tid; i < n; i += 32) {
        double v = fabs(M[(size_t)i * n + k]);
        if (v > best) { best = v; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) piv_s = bi;
    }
    __syncthreads();
    const int p = piv_s;
    if (p != k) {
      for (int j = tid; j < 2 * n; j += kThreads) {
        double* X = (j < n) ? M : B;
        const int jj = (j < n) ? j : j - n;
        double t = X[(size_t)k * n + jj];
        X[(size_t)k * n + jj] = X[(size_t)p * n + jj];
        X[(size_t)p * n + jj] = t;
      }
      __syncthreads();
    }
    const double pinv = 1.0 / M[(size_t)k * n + k];
    // rows i != k, columns k+1..n-1 of M and all columns of B; column k of M is only read.
    // 64 column lanes x 4 row groups.  (The flattened (row, column) loop of the first version
    // spent 70 % of the kernel's instructions here, ~30 per update in 64-bit index arithmetic
    // and integer divisions: ncu source view, 415 k cycles per 61 x 61 matrix.)
    {
      const int tx = tid & 63, ty = tid >> 6;
      if (n <= 64) {
        // one column of M and one of B per thread; row k is read once per step.  Four rows at
        // a time, all loads before the first store (the compiler must assume M and B alias and
        // otherwise serialises load -> fma -> store row by row: 100 cycles per row, latency bound);
        // row k itself takes a zero multiplier instead of a branch.
        const bool cm = tx > k && tx < n, cb = tx < n;
        const double mk = cm ? M[k * n + tx] : 0.0, bk = cb ? B[k * n + tx] : 0.0;
        for (int i0 = ty; i0 < n; i0 += 4 * (kThreads / 64)) {
          double l[4], mv[4], bv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int i = i0 + r * (kThreads / 64);
            const bool ok = i < n;
            const int o = i * n;
            l[r] = (ok && i != k) ? M[o + k] : 0.0;
            mv[r] = (ok && cm) ? M[o + tx] : 0.0;
            bv[r] = (ok && cb) ? B[o + tx] : 0.0;
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int i = i0 + r * (kThreads / 64);
            const bool ok = i < n;
            const int o = i * n;
            const double ll = l[r] * pinv;
            if (ok && cm) M[o + tx] = fma(-ll, mk, mv[r]);
            if (ok && cb) B[o + tx] = fma(-ll, bk, bv[r]);
          }
        }
      } else {
        for (int i = ty; i < n; i += kThreads / 64) {
          if (i == k) continue;
          const double l = M[(size_t)i * n + k] * pinv;
          for (int j = k + 1 + tx; j < n; j += 64)
            M[(size_t)i * n + j] = fma(-l, M[(size_t)k * n + j], M[(size_t)i * n + j]);
          for (int j = tx; j < n; j += 64)
            B[(size_t)i * n + j] = fma(-l, B[(size_t)k * n + j], B[(size_t)i * n + j]);
        }
      }
    }
    __syncthreads();
  }
  for (size_t i = tid; i < nn; i += kThreads) {
    const size_t r = i / n;
    B[i] = B[i] / M[r * n + r];
  }
  __syncthreads();

  // squaring: R = B; ping-pong between B (V) and U
  double* R = B;
  double* O = U;
  for (int i = 0; i < s; ++i) {
    mm(O, R, R, n, sm);
    double* t = R; R = O; O = t;
  }
  for (size_t i = tid; i < nn; i += kThreads) A[i] = R[i];
}

// ---------------------------------------------------------------------------------------
// Small matrices (n <= 16: every S <= 16 transition matrix, the Frechet block form for S <= 8):
// one WARP per matrix, all seven n x n work matrices in shared memory, lanes own elements, only
// __syncwarp between the steps.  The generic kernel above spends ~25 us on a 4 x 4 matrix in
// block-wide barriers around a 64 x 64 tile loop; this one is a dependent chain of ~40 short steps.
// Same algorithm (Pade-13, scaling and squaring, LU with partial pivoting).
// ---------------------------------------------------------------------------------------
constexpr int kSmallMax = 16;
constexpr int kSmallWarps = 4;

__device__ __forceinline__ void mm_w(double* C, const double* A, const double* B, int n, int lane) {
  for (int e = lane; e < n * n; e += 32) {
    const int i = e / n, j = e % n;
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc = fma(A[i * n + k], B[k * n + j], acc);
    C[e] = acc;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kSmallWarps * 32)
expm_small_kernel(double* __restrict__ mats, int n, int n_mat) {
  extern __shared__ double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * kSmallWarps + warp;
  if (m >= n_mat) return;
  const int nn = n * n;
  double* A = smw + (size_t)warp * 8 * nn;
  double *A2 = A + nn, *A4 = A + 2 * nn, *A6 = A + 3 * nn, *T1 = A + 4 * nn, *U = A + 5 * nn,
         *V = A + 6 * nn, *T2 = A + 7 * nn;
  double* G = mats + (size_t)m * nn;
  for (int e = lane; e < nn; e += 32) A[e] = G[e];
  __syncwarp();
  // 1-norm
  double cmax = 0.0;
  for (int j = lane; j < n; j += 32) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += fabs(A[i * n + j]);
    cmax = fmax(cmax, t);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  int s = 0;
  if (cmax > kTheta13) {
    s = (int)ceil(log2(cmax / kTheta13));
    if (s < 0) s = 0;
    const double sc = ldexp(1.0, -s);
    for (int e = lane; e < nn; e += 32) A[e] *= sc;
  }
  __syncwarp();
  const double* b = kPade13;
  mm_w(A2, A, A, n, lane);
  mm_w(A4, A2, A2, n, lane);
  mm_w(A6, A4, A2, n, lane);
  for (int e = lane; e < nn; e += 32) {
    const double a2 = A2[e], a4 = A4[e], a6 = A6[e];
    T1[e] = b[13] * a6 + b[11] * a4 + b[9] * a2;
    T2[e] = b[12] * a6 + b[10] * a4 + b[8] * a2;
  }
  __syncwarp();
  mm_w(U, A6, T1, n, lane);
  mm_w(V, A6, T2, n, lane);
  for (int e = lane; e < nn; e += 32) {
    const double a2 = A2[e], a4 = A4[e], a6 = A6[e];
    const bool diag = (e / n) == (e % n);
    T1[e] = U[e] + b[7] * a6 + b[5] * a4 + b[3] * a2 + (diag ? b[1] : 0.0);
    V[e] = V[e] + b[6] * a6 + b[4] * a4 + b[2] * a2 + (diag ? b[0] : 0.0);
  }
  __syncwarp();
  mm_w(U, A, T1, n, lane);
  for (int e = lane; e < nn; e += 32) {
    const double u = U[e], v = V[e];
    T1[e] = v - u;     // M
    T2[e] = v + u;     // B
  }
  __syncwarp();
  double* M = T1;
  double* B = T2;
  for (int k = 0; k < n; ++k) {
    // pivot: lanes k..n-1 hold |M[i][k]|
    double best = (lane >= k && lane < n) ? fabs(M[lane * n + k]) : -1.0;
    int bi = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    const int p = bi;
    if (p != k) {
      for (int j = lane; j < 2 * n; j += 32) {
        double* X = (j < n) ? M : B;
        const int jj = (j < n) ? j : j - n;
        const double t = X[k * n + jj];
        X[k * n + jj] = X[p * n + jj];
        X[p * n + jj] = t;
      }
    }
    __syncwarp();
    const double pinv = 1.0 / M[k * n + k];
    const int rows = n - k - 1, cols = (n - k - 1) + n;
    // multipliers first (column k is read by every element of its row)
    double lmul = (lane > k && lane < n) ? M[lane * n + k] * pinv : 0.0;
    __syncwarp();
    for (int base = 0; base < rows * cols; base += 32) {     // warp-uniform trip count
      const int idx = base + lane;
      const bool on = idx < rows * cols;
      const int i = on ? k + 1 + idx / cols : 0, c = on ? idx % cols : 0;
      const double l = __shfl_sync(0xffffffffu, lmul, i);
      if (on) {
        if (c < n - k - 1) {
          const int j = k + 1 + c;
          M[i * n + j] = fma(-l, M[k * n + j], M[i * n + j]);
        } else {
          const int j = c - (n - k - 1);
          B[i * n + j] = fma(-l, B[k * n + j], B[i * n + j]);
        }
      }
    }
    __syncwarp();
  }
  for (int j = lane; j < n; j += 32) {
    for (int k = n - 1; k >= 0; --k) {
      double v = B[k * n + j];
      for (int c = k + 1; c < n; ++c) v = fma(-M[k * n + c], B[c * n + j], v);
      B[k * n + j] = v / M[k * n + k];
    }
  }
  __syncwarp();
  double* R = B;
  double* O = U;
  for (int i = 0; i < s; ++i) {
    mm_w(O, R, R, n, lane);
    double* t = R; R = O; O = t;
  }
  for (int e = lane; e < nn; e += 32) G[e] = R[e];
}

static int launch_expm(double* mats, int n, int n_mat, double* scratch, cudaStream_t stream) {
  if (n <= kSmallMax) {
    const size_t smem = sizeof(double) * 8 * (size_t)n * n * kSmallWarps;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(expm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    expm_small_kernel<<<(n_mat + kSmallWarps - 1) / kSmallWarps, kSmallWarps * 32, smem, stream>>>(
        mats, n, n_mat);
  } else {
    const size_t ws = sizeof(double) * 3 * (size_t)n * n;
    if (ws + 20 * 1024 <= 227 * 1024) {
      cudaFuncSetAttribute(expm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws);
      expm_kernel<true><<<n_mat, kThreads, ws, stream>>>(mats, n, scratch);
    } else {
      expm_kernel<false><<<n_mat, kThreads, 0, stream>>>(mats, n, scratch);
    }
  }
  return 0;
}

__global__ void build_scaled_kernel(const double* __restrict__ Q, const int32_t* __restrict__ q_index,
                                    const double* __restrict__ t, int S, double* __restrict__ out) {
  const int m = blockIdx.x;
  const double* Qm = Q + (size_t)(q_index ? q_index[m] : 0) * S * S;
  const double tm = t[m];
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) out[(size_t)m * S * S + i] = Qm[i] * tm;
}

// out[m] = [[tQ^T, c*W_m],[0, tQ^T]] with c = 1/(S*max|W_m|) (0 if W_m == 0), so
// the direction block has 1-norm <= 1 and never drives the scaling;
// scale[m] = 1/c so that L(tQ^T, tW) = t * scale * topright(expm(out)).
__global__ void build_frechet_kernel(const double* __restrict__ Q, const int32_t* __restrict__ q_index,
                                     const double* __restrict__ t, const double* __restrict__ W,
                                     int S, double* __restrict__ out, double* __restrict__ scale) {
  __shared__ double red[32];
  const int m = blockIdx.x;
  const int n = 2 * S;
  const double* Qm = Q + (size_t)(q_index ? q_index[m] : 0) * S * S;
  const double* Wm = W + (size_t)m * S * S;
  const double tm = t[m];
  double mx = 0.0;
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) mx = fmax(mx, fabs(Wm[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.0;
  for (int i = 0; i < (blockDim.x + 31) / 32; ++i) mx = fmax(mx, red[i]);
  const double c = mx > 0.0 ? 1.0 / (mx * S) : 0.0;
  if (threadIdx.x == 0) scale[m] = mx * S;
  double* O = out + (size_t)m * n * n;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx % n;
    double v = 0.0;
    if (i < S && j < S) v = tm * Qm[(size_t)j * S + i];
    else if (i >= S && j >= S) v = tm * Qm[(size_t)(j - S) * S + (i - S)];
    else if (i < S && j >= S) v = c * Wm[(size_t)i * S + (j - S)];
    O[idx] = v;
  }
}

__global__ void extract_frechet_kernel(const double* __restrict__ blk, const double* __restrict__ t,
                                       const double* __restrict__ scale, int S,
                                       double* __restrict__ M) {
  const int m = blockIdx.x;
  const int n = 2 * S;
  const double f = t[m] * scale[m];
  const double* O = blk + (size_t)m * n * n;
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx % S;
    M[(size_t)m * S * S + idx] = f * O[(size_t)i * n + (S + j)];
  }
}

// dwell[c] = sum_{m >= m0} M[m][c][c];  trans[c][d] = sum_{m >= m0} Q_m[c][d] * M[m][c][d] (c != d, 0 on the
// diagonal).  One CTA per entry (c, d), threads stride over the matrices.
// raoteh/sampler/_mjp_dense.py:521-533 (the accumulation over edges).
__global__ void __launch_bounds__(128)
history_stats_kernel(const double* __restrict__ Q, const int32_t* __restrict__ q_index,
                     const double* __restrict__ M, int n_mat, int m0, int S,
                     double* __restrict__ dwell, double* __restrict__ trans) {
  __shared__ double red[4];
  const int idx = blockIdx.x, c = idx / S, d = idx % S;
  double acc = 0.0;
  for (int m = m0 + threadIdx.x; m < n_mat; m += blockDim.x) {
    const double v = M[(size_t)m * S * S + idx];
    acc += (c == d) ? v : Q[(size_t)(q_index ? q_index[m] : 0) * S * S + idx] * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double tot = red[0] + red[1] + red[2] + red[3];
    if (c == d) { dwell[c] = tot; trans[idx] = 0.0; }
    else trans[idx] = tot;
  }
}

// P[m] = A diag(exp(t[m] * lam)) B for every edge m: the spectral form of a time-reversible rate
// matrix Q = S diag(D), A = diag(D^-1/2) U, B = U^T diag(D^1/2), (lam, U) = eigh(D^1/2 S D^1/2)
// (examples/p53/qtop.py:76-85 getp_spectral_v2, :283-288 reconstruct_spectral_v2).  One CTA per
// edge; A scaled by the exponentials and B live in shared memory, a thread owns output entries.
// d_off[i] != 0 marks states with D[i] == 0, whose diagonal entry is set to 1 (qtop.py:83-84).
__global__ void __launch_bounds__(256)
expm_spectral_kernel(const double* __restrict__ A, const double* __restrict__ lam,
                     const double* __restrict__ B, const double* __restrict__ t,
                     const uint8_t* __restrict__ d_off, int S, double* __restrict__ P) {
  extern __shared__ double sms[];
  double* As = sms;                 // [S][S+1]  A[i][k] * exp(t lam_k)
  double* Bs = As + S * (S + 1);    // [S][S]
  const int m = blockIdx.x;
  const double tm = t[m];
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, k = idx % S;
    As[i * (S + 1) + k] = A[idx] * exp(tm * lam[k]);
    Bs[idx] = B[idx];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx % S;
    double acc = 0.0;
    for (int k = 0; k < S; ++k) acc = fma(As[i * (S + 1) + k], Bs[k * S + j], acc);
    if (d_off && i == j && d_off[i]) acc = 1.0;
    P[(size_t)m * S * S + idx] = acc;
  }
}

}  // namespace

// Lower bound of expm(tQ) over a small step (examples/p53/liwen.py:48-82; pyfelscore.
// get_lb_transition_matrix): no change on the diagonal, exactly one change a -> b elsewhere,
//   P[a][a] = exp(t Q[a][a]),  P[a][b] = Q[a][b] (e^{-ra t} - e^{-rb t}) / (rb - ra)   (ra = -Q[a][a]),
//   -> Q[a][b] t e^{-rb t} when ra == rb.
__global__ void lb_transition_kernel(const double* __restrict__ Q, const double* __restrict__ t,
                                     int S, double* __restrict__ P) {
  const double tm = t[blockIdx.x];
  double* Pm = P + (size_t)blockIdx.x * S * S;
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int a = idx / S, b = idx % S;
    double p;
    if (a == b) {
      p = exp(tm * Q[idx]);
    } else {
      const double rab = Q[idx];
      if (rab != 0.0) {
        const double ra = -Q[a * S + a], rb = -Q[b * S + b];
        // (e^{-ra t} - e^{-rb t}) / (rb - ra) without the cancellation of the literal form for
        // small (rb - ra) t:  e^{-ra t} * (-expm1(-(rb - ra) t)) / (rb - ra)
        const double d = rb - ra;
        if (d == 0.0) p = rab * tm * exp(-rb * tm);
        else p = rab * exp(-ra * tm) * (-expm1(-d * tm) / d);
      } else {
        p = 0.0;
      }
    }
    Pm[idx] = p;
  }
}

int rt_lb_transition_impl(const double* Q, const double* t, int n_mat, int S, double* P,
                          cudaStream_t stream) {
  if (n_mat <= 0) return RT_OK;
  lb_transition_kernel<<<n_mat, 256, 0, stream>>>(Q, t, S, P);
  RT_CUDA_CHECK(cudaGetLastError());
  return RT_OK;
}

int rt_expm_spectral_impl(const double* A, const double* lam, const double* B, const double* t,
                          const uint8_t* d_off, int n_mat, int S, double* P, cudaStream_t stream) {
  if (n_mat <= 0) return RT_OK;
  const size_t smem = sizeof(double) * ((size_t)S * (S + 1) + (size_t)S * S);
  if (smem > 48 * 1024)
    RT_CUDA_CHECK(cudaFuncSetAttribute(expm_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  expm_spectral_kernel<<<n_mat, 256, smem, stream>>>(A, lam, B, t, d_off, S, P);
  RT_CUDA_CHECK(cudaGetLastError());
  return RT_OK;
}

// P[m] = expm(Q[q_index[m]] * t[m]) for m in [0, n_mat)
int rt_expm_batched_impl(const double* Q, const int32_t* q_index, const double* t, int n_mat, int S,
                         double* P, cudaStream_t stream) {
  if (n_mat <= 0) return RT_OK;
  if (S < 1 || S > 128) return RT_ERR_UNSUPPORTED;
  double* scratch = nullptr;
  RT_CUDA_CHECK(rt_ws_alloc((void**)&scratch, sizeof(double) * 7 * (size_t)S * S * n_mat, stream));
  build_scaled_kernel<<<n_mat, 256, 0, stream>>>(Q, q_index, t, S, P);
  launch_expm(P, S, n_mat, scratch, stream);
  cudaError_t e = cudaGetLastError();
  rt_ws_free(scratch, stream);
  RT_CUDA_CHECK(e);
  return RT_OK;
}

// M[m] = L(t Q^T, t W[m]) (Frechet derivative of expm), m in [0, n_mat)
int rt_frechet_contract_impl(const double* Q, const int32_t* q_index, const double* t,
                             const double* W, int n_mat, int S, double* M, cudaStream_t stream) {
  if (n_mat <= 0) return RT_OK;
  if (S < 1 || 2 * S > 128) return RT_ERR_UNSUPPORTED;
  const int n = 2 * S;
  double* buf = nullptr;
  const size_t nn = (size_t)n * n;
  RT_CUDA_CHECK(rt_ws_alloc((void**)&buf, sizeof(double) * (8 * nn + 1) * n_mat, stream));
  double* blk = buf;
  double* scratch = buf + nn * n_mat;
  double* scale = buf + 8 * nn * n_mat;
  build_frechet_kernel<<<n_mat, 256, 0, stream>>>(Q, q_index, t, W, S, blk, scale);
  launch_expm(blk, n, n_mat, scratch, stream);
  extract_frechet_kernel<<<n_mat, 256, 0, stream>>>(blk, t, scale, S, M);
  cudaError_t e = cudaGetLastError();
  rt_ws_free(buf, stream);
  RT_CUDA_CHECK(e);
  return RT_OK;
}

int rt_history_statistics_impl(const double* Q, const int32_t* q_index, const double* t,
                               const double* W, int n_mat, int first, int S, double* M,
                               double* dwell, double* trans, cudaStream_t stream) {
  int rc = rt_frechet_contract_impl(Q, q_index, t, W, n_mat, S, M, stream);
  if (rc != RT_OK) return rc;
  if (first > 0) RT_CUDA_CHECK(cudaMemsetAsync(M, 0, sizeof(double) * (size_t)first * S * S, stream));
  history_stats_kernel<<<S * S, 128, 0, stream>>>(Q, q_index, M, n_mat, first, S, dwell, trans);
  RT_CUDA_CHECK(cudaGetLastError());
  return RT_OK;
}
