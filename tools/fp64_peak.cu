// tools/fp64_peak.cu -- measured FP64 peaks of this GPU: DFMA (vector pipe) and DMMA (mma.sync.m8n8k4.f64).
// Denominators of the assembly roofline ("not yet measured").  Not product code.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters) {
  double c[ILP][2];
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_kernel(double* out, int iters) {
  double c[ILP];
  for (int i = 0; i < ILP; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Do DMMA and DFMA share one FP64 datapath?  Even warps run the DMMA loop, odd warps the DFMA loop, in the same block.
// If the two pipes were independent the flops would add up (~70 TFLOP/s); if they are one datapath the mix stays at ~37.
__global__ void mixed_kernel(double* out, int iters) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double c[8];
    for (int i = 0; i < 8; ++i) c[i] = i;
    const double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-6;
    for (int it = 0; it < iters * 8; ++it) {  // 8 x 32 FMA per DMMA-equivalent of 256 FMA: same flops per iteration as below
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
    }
    for (int i = 0; i < 8; ++i) s += c[i];
  } else {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float run(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 3; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  double* out; cudaMalloc(&out, 148 * 32 * 1024 * 8);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    const int blocks = 148 * 2, threads = warps * 32 / 2 < 32 ? 32 : warps * 32 / 2;
    float t = run([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters); });
    double flops = (double)blocks * (threads / 32) * iters * 8 * 512.0;
    printf("DMMA m8n8k4  %2d warps/SM ILP8: %.3f ms  %.2f TFLOP/s\n", warps, t, flops / t / 1e9);
  }
  for (int warps : {8, 16, 32}) {
    const int blocks = 148 * 2, threads = warps * 32 / 2;
    float t = run([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters); });
    double flops = (double)blocks * (threads / 32) * iters * 16 * 512.0;
    printf("DMMA m8n8k4  %2d warps/SM ILP16: %.3f ms  %.2f TFLOP/s\n", warps, t, flops / t / 1e9);
  }
  for (int warps : {8, 16, 32, 64}) {
    const int blocks = 148 * 2, threads = warps * 32 / 2;
    float t = run([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters); });
    double flops = (double)blocks * threads * iters * 8 * 2.0;
    printf("DFMA         %2d warps/SM ILP8: %.3f ms  %.2f TFLOP/s\n", warps, t, flops / t / 1e9);
  }
  {
    // 16 warps / SM: 8 DMMA warps + 8 DFMA warps, each issuing iters * 8 * 512 flops
    const int blocks = 148 * 2, threads = 8 * 32;
    float t = run([&] { mixed_kernel<<<blocks, threads>>>(out, iters / 4); });
    double flops = (double)blocks * (threads / 32) * (iters / 4) * 8 * 512.0;
    printf("MIXED 8 DMMA + 8 DFMA warps/SM: %.3f ms  %.2f TFLOP/s (sum of both halves)\n", t, flops / t / 1e9);
  }
  return 0;
}
