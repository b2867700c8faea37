// FP64 peak probe for the roofline denominators of the 61-state kernels:
// measures DMMA (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4) and DFMA issue
// throughput on all SMs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o fp64_peak tools/fp64_peak.cu ; run on the B200 (tools/run_probe.sh).
#include <cuda_runtime.h>
#include <stdio.h>

template <int ILP>
__global__ void dmma_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      dmma_kernel<8><<<sms, warps * 32>>>(out, iters);
    }
    cudaEventRecord(e0);
    dmma_kernel<8><<<sms, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)sms * warps * iters * 8 * 512.0;
    printf("{\"probe\":\"dmma_m8n8k4\",\"warps_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", warps, flops / ms * 1e-9, ms);
    cudaEventRecord(e0);
    dfma_kernel<8><<<sms, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    flops = (double)sms * warps * 32 * iters * 8 * 2.0;
    printf("{\"probe\":\"dfma\",\"warps_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", warps, flops / ms * 1e-9, ms);
  }
  printf("{\"sms\":%d,\"name\":\"%s\",\"clock_khz\":%d}\n", sms, p.name, p.clockRate);
  return 0;
}
